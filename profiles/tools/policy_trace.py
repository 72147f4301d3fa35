#!/usr/bin/env python
"""Timeline of one tile of the fused policy kernel (block 0, second tile, actor tower): clock64 stamps of the epilogue warps
(accumulator ready -> item done) and of the MMA issuer (dependencies met -> committed), printed relative to the tile's first stamp."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200 import _capi  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.fused_policy import FusedPolicy  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.ppo import QuadActorCritic  # noqa: E402

dev = torch.device("cuda:0")
pol = QuadActorCritic(QuadSimConfig(num_envs=8, num_agents=8)).to(dev)
fp = FusedPolicy(pol, dev)
grid = int(os.environ.get("QP_MAX_GRID", "148"))
obs = torch.randn(grid * 128 * 4, 54, device=dev)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
L = _capi.lib()
L.qp_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
fp.forward(obs)
L.qp_debug_trace(fp._h, buf.data_ptr())
fp.forward(obs)
torch.cuda.synchronize()
L.qp_debug_trace(fp._h, None)
b = buf.cpu().tolist()
epi = [x for x in b[:256] if x]
mma = [x for x in b[256:] if x]
t0 = min(epi + mma)
names = []
V = fp.V
for j0 in range(0, V, 2):
    nact = min(2, V - j0)
    for it in range(4):
        for s in range(nact):
            names.append(f"nbr{j0 + s} L{1 + it // 2}h{it & 1} s{s}")
names += ["self L1h0 s0", "self L1h1 s1", "self L2h0 s0", "self L2h1 s1"] + [f"ff q{q} s{q & 1}" for q in range(4)]
print(f"{len(epi)} epilogue stamps, {len(mma)} MMA stamps, tile span {max(epi + mma) - t0} clk")
print("EPILOGUE (ready, done, duration; gap = wait since the previous item finished)")
prev = None
for i in range(0, len(epi) - 1, 2):
    print(f"  {epi[i] - t0:7d} {epi[i + 1] - t0:7d}  dur {epi[i + 1] - epi[i]:6d}  gap {'' if prev is None else epi[i] - prev:>6}")
    prev = epi[i + 1]
print("MMA issuer stamps (relative)")
print("  " + " ".join(str(x - t0) for x in mma))
print("items in order: " + ", ".join(names))
