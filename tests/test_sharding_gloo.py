"""World-size-2 test of the multi-GPU host logic on CPU (gloo): env sharding keyed by global env id and the episode-stat
all-reduce.  Each rank steps its shard (oracle-backed simulator); together they must reproduce the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total_envs, steps, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle_sim import OracleSim
    from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig
    from quad_swarm_rl_stable_baselines3_b200.sharding import all_reduce_stats, shard_config
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = QuadSimConfig(num_envs=total_envs, num_agents=4, ep_time=0.08, seed=7, neighbor_visible_num=2)
    mine = shard_config(cfg, rank, world)
    sim = OracleSim(mine)
    K = cfg.num_agents
    lo = mine.env_id_offset
    rs = np.random.RandomState(5)
    acts = rs.uniform(-1, 1, (steps, total_envs * K, 4)).astype(np.float32)       # same global action tensor on every rank
    obs = [sim.reset_host()]
    for t in range(steps):
        o, r, d = sim.step_host(acts[t, lo * K:(lo + mine.num_envs) * K])
        obs.append(o.copy())
    stats = all_reduce_stats(sim.episode_stats())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs=np.stack(obs), lo=lo, n=mine.num_envs,
             episodes=stats["episodes"], collisions=stats["num_collisions"])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from quad_swarm_rl_stable_baselines3_b200.sharding import shard_range
    for total, world in ((65536, 8), (10, 3), (7, 7), (4097, 4)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(2, 0, 3)


@pytest.mark.timeout(300)
def test_two_rank_run_reproduces_single_rank(tmp_path):
    total_envs, steps, world = 5, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total_envs, steps, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_sim import OracleSim
    from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig
    cfg = QuadSimConfig(num_envs=total_envs, num_agents=4, ep_time=0.08, seed=7, neighbor_visible_num=2)
    sim = OracleSim(cfg)
    rs = np.random.RandomState(5)
    acts = rs.uniform(-1, 1, (steps, total_envs * 4, 4)).astype(np.float32)
    ref = [sim.reset_host()]
    for t in range(steps):
        ref.append(sim.step_host(acts[t])[0].copy())
    ref = np.stack(ref)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [0, 3] and [int(p["n"]) for p in parts] == [3, 2]
    got = np.concatenate([p["obs"] for p in parts], axis=1)
    np.testing.assert_array_equal(got, ref)                    # shards == the single-process run, bit for bit
    tot = sim.episode_stats()
    assert tot["episodes"] >= total_envs
    for p in parts:                                            # every rank holds the global sums after the all-reduce
        assert int(p["episodes"]) == tot["episodes"] and int(p["collisions"]) == tot["num_collisions"]
