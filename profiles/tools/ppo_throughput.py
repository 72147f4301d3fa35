#!/usr/bin/env python
"""BASELINE.json configs[4] "with PPO update": rollout (policy inference + simulator) and update throughput of the device
PPO loop.  Single GPU: python ppo_throughput.py [envs] [n_steps]; N GPUs: torchrun --nproc-per-node N ppo_throughput.py ..."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.ppo import DevicePPO, PPOConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim  # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = QuadSimConfig(num_envs=envs, num_agents=8, seed=0, env_id_offset=rank * envs)
sim = QuadSwarmSim(cfg, device=dev)
sim.want_terminal_obs = False
ppo = DevicePPO(sim, cfg, PPOConfig(n_steps=n_steps, batch_size=65536, n_epochs=1, autocast_bf16=True))
ppo.learn(1)                                                    # warm-up (cuBLAS heuristics, allocator)
hist = ppo.learn(2)
n = envs * 8
roll = sum(r["rollout_s"] for r in hist) / len(hist)
upd = sum(r["update_s"] for r in hist) / len(hist)
t = torch.tensor([roll, upd], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    roll, upd = float(t[0]), float(t[1])
    print(json.dumps({"n_gpus": world, "envs_per_gpu": envs, "agents": 8, "n_steps": n_steps,
                      "rollout_drone_steps_per_s": world * n * n_steps / roll, "update_samples_per_s": world * n * n_steps / upd,
                      "train_drone_steps_per_s": world * n * n_steps / (roll + upd), "rollout_s": roll, "update_s": upd,
                      "policy": "2 towers x (self 18-256-256, deep-sets neighbours 24-256-256, ff 512-512), bf16 autocast"}))
if world > 1:
    dist.destroy_process_group()
