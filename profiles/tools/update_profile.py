import sys, os, torch
sys.path.insert(0, "/root/repo")
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig
from quad_swarm_rl_stable_baselines3_b200.ppo import DevicePPO, PPOConfig
from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
cfg = QuadSimConfig(num_envs=16384, num_agents=8, seed=0)
sim = QuadSwarmSim(cfg, device="cuda:0"); sim.want_terminal_obs = False
ppo = DevicePPO(sim, cfg, PPOConfig(n_steps=2, batch_size=65536, n_epochs=1, autocast_bf16=True))
ppo.collect(); ppo.update()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ppo.update()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
