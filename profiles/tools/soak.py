#!/usr/bin/env python
"""Soak run: many control steps of every mode at the BASELINE batch size with random actions; prints the episode aggregate and
fails on a non-finite observation / reward or a non-finite-state reset.  usage: soak.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = 65536
cases = {
    "cfg2": QuadSimConfig(num_envs=n, num_agents=8, seed=11),
    "cfg3": QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                          obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, seed=12),
    "mix": QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", seed=13),
    "cfg4": QuadSimConfig(num_envs=n // 4, num_agents=32, seed=14),
    "fork": QuadSimConfig.fork_default(num_envs=n, seed=15),
}
for name, cfg in cases.items():
    sim = QuadSwarmSim(cfg, device="cuda:0")
    sim.want_terminal_obs = False
    g = torch.Generator(device="cuda").manual_seed(1)
    pool = torch.rand((16, cfg.num_envs * cfg.num_agents, cfg.act_dim), device="cuda", generator=g) * 2 - 1
    pool[8:] = pool[8:] * 0.15 + 0.05                      # half of the time near hover: long flights, formations reached
    sim.reset()
    k = steps // (8 if name == "fork" else 1)
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for s in range(k):
        obs, rew, done = sim.step(pool[(s // 200) % 16])
        if s % 97 == 0:
            bad += (~torch.isfinite(obs)).sum() + (~torch.isfinite(rew)).sum()
    st = sim.episode_stats()
    print(name, "steps", k, "episodes", st["episodes"], "nonfinite_resets", st["nonfinite_resets"], "nonfinite values", int(bad),
          "collisions/episode %.2f" % (st["num_collisions"] / max(st["episodes"], 1)),
          "success/deadlock/collided per episode %.2f %.2f %.2f" % tuple(st[k2] / max(st["episodes"], 1) for k2 in ("agents_success", "agents_deadlock", "agents_collided")))
    assert int(bad) == 0 and st["nonfinite_resets"] == 0 and st["episodes"] > 0
    del sim
print("soak ok")
