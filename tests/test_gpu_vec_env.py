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


@pytest.mark.parametrize("mode", ["upstream", "mix", "obstacles", "fork"])
def test_host_path_chunked_pipeline_is_bitwise_the_single_launch(mode, monkeypatch):
    """qs_step_host cuts large batches into env chunks on two internal streams (D2H of one chunk overlaps the H2D + kernel of the
    next).  Environments are independent and the RNG is keyed by the global env id, so the chunked step must equal the single
    launch bit for bit -- observations, rewards, dones, terminal observations, reset_infos and the episode records."""
    import torch
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
    kw = dict(num_envs=200, num_agents=8, ep_time=0.06, seed=5)
    if mode == "mix":
        cfg = QuadSimConfig(quads_mode="mix", **kw)
    elif mode == "obstacles":
        cfg = QuadSimConfig(quads_mode="mix", use_obstacles=True, use_downwash=True, obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, **kw)
    elif mode == "fork":
        cfg = QuadSimConfig.fork_default(num_envs=200, num_agents=4, ep_time=0.3, capture_radius=2.5, seed=5)
    else:
        cfg = QuadSimConfig(**kw)
    monkeypatch.setenv("QS_HOST_CHUNKS", "1")
    a_sim = QuadSwarmSim(cfg, device="cuda:0")
    monkeypatch.setenv("QS_HOST_CHUNKS", "3")                       # 200 envs -> chunks of 96, 96, 8
    b_sim = QuadSwarmSim(cfg, device="cuda:0")
    n, D, A = cfg.num_envs * cfg.num_agents, cfg.obs_dim, cfg.act_dim
    assert np.array_equal(a_sim.reset_host(), b_sim.reset_host())
    rs = np.random.RandomState(1)
    dones = 0
    for t in range(12):
        act = rs.uniform(-1, 1, (n, A)).astype(np.float32)
        outs = []
        for k, sim in enumerate((a_sim, b_sim)):
            monkeypatch.setenv("QS_HOST_CHUNKS", "1" if k == 0 else "3")
            term, succ = np.zeros((n, D), np.float32), np.zeros(cfg.num_envs, np.uint8)
            obs, rew, done = sim.step_host(act, terminal_obs=term, reset_success=succ)
            outs.append((obs.copy(), rew.copy(), np.asarray(done).copy(), term, succ, *sim.episode_records_host()))
        for x, y in zip(*outs):
            assert np.array_equal(x, y, equal_nan=True)
        dones += int(outs[0][2].any())
    assert dones >= 1
    sa, sb = a_sim.get_state(), b_sim.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert a_sim.episode_stats()["episodes"] == b_sim.episode_stats()["episodes"] > 0
