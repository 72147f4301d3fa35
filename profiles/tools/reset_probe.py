#!/usr/bin/env python
"""Where does the steady-state step time go?  For each workload: (a) the explicit reset kernel on a partial wave (= the latency of
one warp's reset path, which is what an in-step auto-reset adds to the warp that hosts it), (b) the step at the bench's steady
state (staggered clocks, ~N/ep_len resets per step), (c) the same landed-drone mix with episodes that never end (no resets in the
window).  (b) - (c) is what the resets cost; usage: reset_probe.py [case ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim  # noqa: E402


def cfg_of(case, n, **kw):
    if case == "cfg2":
        return QuadSimConfig(num_envs=n, num_agents=8, **kw)
    if case == "mix":
        return QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", **kw)
    if case == "cfg3":
        return QuadSimConfig(num_envs=n, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                             obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, **kw)
    if case == "cfg4":
        return QuadSimConfig(num_envs=n // 4, num_agents=32, **kw)
    raise KeyError(case)


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for case in (sys.argv[1:] or ["cfg2", "cfg3", "mix", "cfg4"]):
    g = torch.Generator(device="cuda").manual_seed(1)
    # (a) reset kernel, partial wave
    small = QuadSwarmSim(cfg_of(case, 2048), device="cuda:0")
    for _ in range(5):
        small.reset()
    t_reset = min(timed(lambda i: small.reset(), 20) for _ in range(3))
    small.close()
    out = [f"{case}: reset kernel @2048 envs {t_reset:7.2f} us"]
    for label, kw, stagger in (("steady", {}, True), ("no-reset", dict(ep_time=1.0e5), False)):
        cfg = cfg_of(case, 65536, **kw)
        sim = QuadSwarmSim(cfg, device="cuda:0")
        sim.want_terminal_obs = False
        pool = torch.rand((8, cfg.num_envs * cfg.num_agents, 4), device="cuda", generator=g) * 2 - 1
        sim.reset()
        if stagger:
            sim.set_state(tick=torch.randint(0, cfg.ep_len, (cfg.num_envs,), generator=g, device="cuda", dtype=torch.int32))
        for i in range(1564):
            sim.step(pool[i % 8])
        t = min(timed(lambda i: sim.step(pool[i % 8]), 300) for _ in range(3))
        out.append(f"{label} {t:7.2f} us")
        sim.close()
    print(" | ".join(out), flush=True)
