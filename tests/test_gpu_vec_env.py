"""The VecEnv facade over the real CUDA simulator (SubprocVecEnvCustom's surface), fork and upstream modes."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.vec_env import QuadSwarmVecEnv  # noqa: E402


def test_fork_vec_env_on_gpu_matches_oracle_backed_facade():
    """Same seeds and actions through the facade over the CUDA library and over the CPU oracle: same dones, same
    reset_infos, observations within the fork tolerance (free-running for a few calls, no teacher forcing)."""
    from oracle_sim import OracleSim
    cfg = QuadSimConfig.fork_default(num_envs=24, num_agents=4, ep_time=0.16, capture_radius=2.5, seed=21)
    gpu = QuadSwarmVecEnv(cfg, device="cuda:0")
    cpu = QuadSwarmVecEnv(cfg, sim=OracleSim(cfg))
    o1, o2 = gpu.reset(), cpu.reset()
    np.testing.assert_allclose(o1, o2, atol=5e-4)
    rs = np.random.RandomState(3)
    n_reset = 0
    for t in range(5):
        a = rs.uniform(-1, 1, (gpu.num_envs, 2)).astype(np.float32)
        r1, r2 = gpu.step(a), cpu.step(a)
        assert np.array_equal(r1[2], r2[2])
        np.testing.assert_allclose(r1[1], r2[1], atol=1e-4)
        np.testing.assert_allclose(r1[0], r2[0], atol=2e-3)
        assert gpu.reset_infos == cpu.reset_infos
        for i1, i2 in zip(r1[3], r2[3]):
            assert set(i1) == set(i2)
            if "terminal_observation" in i1:
                n_reset += 1
                np.testing.assert_allclose(i1["terminal_observation"], i2["terminal_observation"], atol=2e-3)
                e1, e2 = i1["episode_extra_stats"], i2["episode_extra_stats"]      # quadrotor_multi_rewards.py:886-978
                assert set(e1) == set(e2) and "dynamic_repulsive/agent_col_rate" in e1
                for k in e1:
                    assert (np.isnan(e1[k]) and np.isnan(e2[k])) or abs(e1[k] - e2[k]) < 1e-6, (k, e1[k], e2[k])
    assert n_reset > 0
    gpu.env_method("set_capture_radius", 0.4)
    assert gpu.get_attr("capture_radius") == [0.4] * cfg.num_envs
    gpu.close()


def test_upstream_vec_env_tensor_and_host_faces_agree():
    import torch
    cfg = QuadSimConfig(num_envs=32, num_agents=8, ep_time=0.06, seed=2)
    a_env, b_env = QuadSwarmVecEnv(cfg, device="cuda:0"), QuadSwarmVecEnv(cfg, device="cuda:0")
    o_host = a_env.reset()
    o_dev = b_env.reset_tensor()
    assert np.array_equal(o_host, o_dev.cpu().numpy())
    rs = np.random.RandomState(0)
    dones = 0
    for t in range(10):
        a = rs.uniform(-1, 1, (a_env.num_envs, 4)).astype(np.float32)
        obs, rew, done, infos = a_env.step(a)
        t_obs, t_rew, t_done = b_env.step_tensor(torch.from_numpy(a).cuda())
        assert np.array_equal(obs, t_obs.cpu().numpy()) and np.array_equal(rew, t_rew.cpu().numpy())
        assert np.array_equal(done, t_done.cpu().numpy())
        if done.any():
            dones += 1
            rows = np.flatnonzero(done)
            term = b_env.sim.terminal_obs.cpu().numpy()
            for r in rows[:8]:
                assert np.array_equal(infos[r]["terminal_observation"], term[r])
                es = infos[r]["episode_extra_stats"]
                assert es["num_collisions"] >= es["num_collisions_after_settle"] >= 0 and es["distance_to_goal_1s"] > 0
                assert abs(es["metric/agent_success_rate"] + es["metric/agent_deadlock_rate"] + es["metric/agent_col_rate"] - 1.0) < 1e-6
    assert dones >= 1
    st = a_env.sim.episode_stats(reduce=True)
    assert st["episodes"] >= cfg.num_envs
