#!/usr/bin/env python
"""One machine, one core: the reference's own step (gym_art/quadrotor_multi, numba JIT on, as sb_train.py / SubprocVecEnvCustom
workers run it) next to the C oracle port that `bench.py`'s CPU arm times.  The reference is Python and cannot travel to the GPU
box, so this factor is measured where /root/reference exists (the build container) and recorded in profiles/ and DESIGN.md:
    reference drone-steps/s/core  x  factor  =  port drone-steps/s/core.
usage: python profiles/tools/ref_vs_port.py [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
out = {}
for name, kw in (("cfg2 (8 quads, static_same_goal, obs 54)", dict(num_agents=8)),
                 ("cfg3 (8 quads, obstacles, obs 40)", dict(num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                                          obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2))):
    import contextlib
    import io
    import ref_harness as rh
    with contextlib.redirect_stdout(io.StringIO()):
        env = rh.make_upstream_env(use_numba=True, **kw)
        env.reset()
    rs = np.random.RandomState(0)
    K = kw["num_agents"]
    acts = rs.uniform(-1, 1, (steps + 50, K, 4))
    with contextlib.redirect_stdout(io.StringIO()):
        for t in range(50):                          # numba compilation + warm-up
            env.step(list(acts[t]))
        t0 = time.perf_counter()
        for t in range(50, 50 + steps):
            o, r, d, i = env.step(list(acts[t]))
        dt_ref = time.perf_counter() - t0
    from oracle import OracleBatch  # noqa: E402
    from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
    ck = dict(kw)
    cfg = QuadSimConfig(num_envs=64, seed=1, **ck)
    ob = OracleBatch(cfg, threads=1)
    ob.reset()
    a = rs.uniform(-1, 1, (64 * K, 4))
    for _ in range(20):
        ob.step(a)
    n_port = 400
    t0 = time.perf_counter()
    for _ in range(n_port):
        ob.step(a)
    dt_port = time.perf_counter() - t0
    ref_v, port_v = steps * K / dt_ref, n_port * 64 * K / dt_port
    out[name] = {"reference_drone_steps_per_s_per_core": ref_v, "port_drone_steps_per_s_per_core": port_v, "port_over_reference": port_v / ref_v,
                 "reference_steps_timed": steps, "port_env_steps_timed": n_port * 64}
out["machine"] = {"cpus": os.cpu_count(), "note": "build container, 1 thread each; numba JIT on for the reference (use_numba=True)"}
print(json.dumps(out, indent=1))
