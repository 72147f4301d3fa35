"""Multi-GPU plumbing: environments shard across ranks with no collective in step/reset.

The reference parallelises over OS processes, one env each (swarm_rl/env_wrappers/subproc_vec_env_custom.py:112-134);
here rank g of G owns a contiguous env range and keys its random streams by GLOBAL env id (`env_id_offset`), so a
G-rank run reproduces the single-rank run of the same total size.  The only collective is the reduction of the small
episode-stat vector (the fields of `episode_extra_stats`, gym_art/quadrotor_multi/quadrotor_multi.py:741-831).
"""
from __future__ import annotations

import copy
from typing import Dict, Tuple

from .config import QsStatsC, QuadSimConfig


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the envs rank `rank` owns: contiguous, sizes differ by at most one, earlier ranks get the extras."""
    if not (0 <= rank < world) or total_envs < world:
        raise ValueError("need 0 <= rank < world <= total_envs")
    base, extra = divmod(total_envs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_config(cfg: QuadSimConfig, rank: int, world: int) -> QuadSimConfig:
    """This rank's slice of a `cfg.num_envs`-env job (strong split).  For weak scaling pass a cfg whose num_envs is
    already world * per-rank."""
    lo, hi = shard_range(cfg.num_envs, rank, world)
    out = copy.copy(cfg)
    out.num_envs = hi - lo
    out.env_id_offset = cfg.env_id_offset + lo
    return out


STAT_FIELDS = tuple(n for n, _ in QsStatsC._fields_)


def all_reduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """Sum the episode statistics of all ranks (torch.distributed: nccl with CUDA tensors, gloo with CPU tensors)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    t = torch.tensor([float(stats[k]) for k in STAT_FIELDS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = t.tolist()
    return {k: (int(round(v)) if QsStatsC._fields_[i][1].__name__ == "c_long" else v) for i, (k, v) in enumerate(zip(STAT_FIELDS, vals))}
