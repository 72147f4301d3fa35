#!/usr/bin/env python
"""Run a few steps of one named workload (profiling aid for ncu): run_case.py <cfg2|cfg3|cfg4|fork|mix> [envs] [steps] [steady]
`steady`: bring the batch to bench.py's steady state first (episode clocks staggered uniformly + more than one episode of random-action
pre-roll), so that the captured launch contains floor contact, collisions and auto-resets like the timed window of the bench."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim  # noqa: E402

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cfg = {"cfg2": lambda: QuadSimConfig(num_envs=n, num_agents=8),
       "cfg3": lambda: QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                     obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2),
       "cfg4": lambda: QuadSimConfig(num_envs=n, num_agents=32),
       "mix": lambda: QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix"),
       "fork": lambda: QuadSimConfig.fork_default(num_envs=n)}[case]()
sim = QuadSwarmSim(cfg, device="cuda:0")
sim.want_terminal_obs = False
a = torch.rand((cfg.num_envs * cfg.num_agents, cfg.act_dim), device="cuda") * 2 - 1
sim.reset()
if len(sys.argv) > 4 and sys.argv[4] == "steady":
    g = torch.Generator(device="cuda").manual_seed(1234)
    pool = torch.rand((8, cfg.num_envs * cfg.num_agents, cfg.act_dim), device="cuda", generator=g) * 2 - 1
    sim.set_state(tick=torch.randint(0, cfg.ep_len, (cfg.num_envs,), generator=g, device="cuda", dtype=torch.int32))
    calls = cfg.ep_len // (cfg.fork.substeps if cfg.env_mode == "fork" else 1) + 64
    for i in range(calls):
        sim.step(pool[i % 8])
    torch.cuda.synchronize()
for _ in range(steps):
    sim.step(a)
torch.cuda.synchronize()
print("ok", case, n, steps)
