#!/usr/bin/env python
"""Fused tcgen05 policy forward vs the torch module it replaces, on the rollout batch of BASELINE configs[4] (65536 envs x 8 quads =
524288 observation rows, obs 54): device time per forward (CUDA events on the launching stream, median of `reps` after warm-up),
achieved dense TFLOP/s against MEASURED_PEAKS.json.  python policy_bench.py [rows] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.fused_policy import FusedPolicy  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.ppo import QuadActorCritic  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
cfg = QuadSimConfig(num_envs=8, num_agents=8)
torch.manual_seed(0)
pol = QuadActorCritic(cfg).to(dev)
fp = FusedPolicy(pol, dev)
obs = torch.randn(rows, 54, device=dev)
S, W, V, A = fp.S, fp.W, fp.V, fp.A
flop_row = 2 * 2 * (S * 256 + 256 * 256 + V * (W * 256 + 256 * 256) + 512 * 512 + 512 * (A + 1) / 2)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def eager_fp32():
    with torch.no_grad():
        return pol.action_net(pol.actor(obs)), pol.value(obs)


def eager_bf16():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return pol.action_net(pol.actor(obs)), pol.value(obs)


peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
out = {"rows": rows, "flop_per_row": flop_row, "policy": f"2 towers x (self {S}-256-256, deep-sets {W}-256-256 x {V}, ff 512-512, heads {A}/1)"}
for name, fn in (("fused_tcgen05", lambda: fp.forward(obs)), ("torch_bf16_autocast", eager_bf16), ("torch_fp32", eager_fp32)):
    med, best = timed(fn)
    out[name] = {"ms_median": med, "ms_min": best, "rows_per_s": rows / med * 1e3, "tflops": flop_row * rows / med * 1e-9,
                 "frac_of_sustained_bf16_peak": flop_row * rows / med * 1e-9 / peaks["bf16_tflops_sustained"]}
m, v = fp.forward(obs)
rm, rv = eager_fp32()
out["max_abs_diff_vs_fp32"] = float(torch.maximum((m - rm).abs().max(), (v - rv).abs().max()))
print(json.dumps(out))
