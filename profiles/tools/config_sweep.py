#!/usr/bin/env python
"""Device-resident step time of every BASELINE.json config shape and a size sweep (informational; bench.py is the bench).
Steps are captured in a CUDA graph (20 per replay) so that small batches are not host-launch-bound."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim  # noqa: E402

CASES = []
for n in (1024, 4096, 16384, 65536, 262144):
    CASES.append((f"cfg2 K=8 obs54, {n} envs", lambda n=n: QuadSimConfig(num_envs=n, num_agents=8), 497.0))
for n in (4096, 65536):
    CASES.append((f"cfg3 K=8 obstacles+downwash obs40, {n} envs",
                  lambda n=n: QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                            obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2), 453.0))
for n in (1024, 16384):
    CASES.append((f"cfg4 K=32 obs54, {n} envs", lambda n=n: QuadSimConfig(num_envs=n, num_agents=32), 501.0))
for n in (4096, 65536):
    CASES.append((f"fork K=4 (8 control steps per call), {n} envs", lambda n=n: QuadSimConfig.fork_default(num_envs=n), 537.0))

dev = torch.device("cuda", 0)
rows = []
for name, mk, bytes_per in CASES:
    cfg = mk()
    sim = QuadSwarmSim(cfg, device=dev)
    sim.want_terminal_obs = False
    nd = cfg.num_envs * cfg.num_agents
    acts = torch.rand((4, nd, cfg.act_dim), device=dev) * 2 - 1
    sim.reset()
    for i in range(30):
        sim.step(acts[i % 4])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20):
            sim.step(acts[i % 4])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 20)
    per_call = cfg.fork.substeps if cfg.env_mode == "fork" else 1
    rows.append(dict(case=name, us_per_call=us, drone_steps_per_s=nd * per_call / (us * 1e-6),
                     hbm_frac=bytes_per * nd / (us * 1e-6) / 6552.6e9))
    del sim
for r in rows:
    print(f"| {r['case']} | {r['us_per_call']:.1f} | {r['drone_steps_per_s']:.3e} | {r['hbm_frac']:.3f} |")
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
