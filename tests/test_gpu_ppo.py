"""Device PPO loop over the real CUDA simulator: rollouts and updates stay on the GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402


@pytest.mark.parametrize("make", [
    lambda: QuadSimConfig(num_envs=256, num_agents=8, ep_time=0.3, seed=1),
    lambda: QuadSimConfig(num_envs=128, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                          obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, ep_time=0.3, seed=2),
    lambda: QuadSimConfig.fork_default(num_envs=256, ep_time=1.0, seed=3),
])
def test_ppo_iterations_on_device(make):
    import torch
    from quad_swarm_rl_stable_baselines3_b200.ppo import DevicePPO, PPOConfig
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
    cfg = make()
    sim = QuadSwarmSim(cfg, device="cuda:0")
    sim.want_terminal_obs = False
    ppo = DevicePPO(sim, cfg, PPOConfig(n_steps=16, batch_size=4096, n_epochs=2, hidden=64, neighbor_hidden=32))
    w0 = torch.cat([p.detach().reshape(-1).clone() for p in ppo.policy.parameters()])
    l0 = sim.launch_count
    hist = ppo.learn(3)
    n = cfg.num_envs * cfg.num_agents
    assert hist[-1]["agent_steps"] == 3 * 16 * n
    assert sim.launch_count - l0 == 3 * 16                      # one simulator launch per env step, nothing else
    for r in hist:
        assert all(np.isfinite(r[k]) for k in ("pg", "vf", "ent", "kl", "mean_reward"))
        assert r["minibatches"] == 2 * int(np.ceil(16 * n / 4096))
    w1 = torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()])
    assert bool(torch.isfinite(w1).all()) and not torch.equal(w0, w1)
    assert ppo.obs_buf.is_cuda and ppo.adv.is_cuda
    assert sum(r["episodes"] for r in hist) > 0

