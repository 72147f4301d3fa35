"""GPU: the per-step reward breakdown (`qs_set_reward_info`, infos[i]["rewards"] / ["goal_dist"]) against the oracle, and the
QuadrotorEnvMulti-shaped facade on the CUDA simulator."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import OracleEnv  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from test_gpu_parity import action_batch, push_state  # noqa: E402


def _sim(cfg):
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
    return QuadSwarmSim(cfg, device="cuda:0")


@pytest.mark.parametrize("kw", [
    dict(num_envs=24, num_agents=8, room_dims=(4.0, 4.0, 6.0), ep_time=0.4),
    dict(num_envs=16, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True, obs_repr="xyz_vxyz_R_omega_floor",
         neighbor_visible_num=2, ep_time=0.4)])
def test_reward_info_matches_oracle(kw):
    cfg = QuadSimConfig(seed=31, **kw)
    sim = _sim(cfg)
    ri = sim.enable_reward_info()
    assert ri.shape == (cfg.num_envs * cfg.num_agents, 8)
    oracles = [OracleEnv(cfg, i) for i in range(cfg.num_envs)]
    sim.reset()
    for o in oracles:
        o.reset()
    rs = np.random.RandomState(1)
    K, N = cfg.num_agents, cfg.num_envs
    hits = np.zeros(8)
    for s in range(50):
        push_state(sim, oracles, cfg)
        a = action_batch(rs, N * K, "hover" if cfg.use_obstacles else "high")
        obs, rew, done = sim.step(torch.from_numpy(a).cuda())
        ref = []
        for e, o in enumerate(oracles):
            o.step(a[e * K:(e + 1) * K].astype(np.float64))
            ref.append(o.reward_info())
        ref = np.concatenate(ref)
        got = sim.reward_info_host()
        # discrete columns (collision / obstacle hit) can differ at a threshold tie: compare rows whose flags agree
        same = (got[:, 5] == ref[:, 5]) & (got[:, 7] == ref[:, 7])
        assert same.mean() > 0.99
        np.testing.assert_allclose(got[same][:, :5], ref[same][:, :5], rtol=2e-5, atol=2e-7)
        np.testing.assert_allclose(got[same][:, 6], ref[same][:, 6], rtol=1e-4, atol=2e-6)
        hits += (ref != 0).sum(axis=0)
        # the weighted sum of the parts is the reward the step returned
        c = cfg.to_c()
        total = (c.rew_pos * got[:, 0] + c.rew_effort * got[:, 1] + c.rew_crash * got[:, 2] + c.rew_orient * got[:, 3] + c.rew_spin * got[:, 4] +
                 c.rew_quadcol_bin * got[:, 5] + got[:, 6] + c.rew_quadcol_bin_obst * got[:, 7])
        np.testing.assert_allclose(total, rew.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert hits[0] > 0 and hits[4] > 0
    if not cfg.use_obstacles:
        assert hits[5] > 0 and hits[6] > 0, "the crowded room must produce collisions and proximity penalties"
    sim.enable_reward_info(False)
    sim.step(torch.zeros((N * K, 4), device="cuda"))


def test_fork_goal_dist_matches_oracle():
    cfg = QuadSimConfig.fork_default(num_envs=12, num_agents=4, ep_time=0.6, capture_radius=2.4, seed=9)
    sim = _sim(cfg)
    sim.enable_reward_info()
    oracles = [OracleEnv(cfg, i) for i in range(cfg.num_envs)]
    sim.reset()
    for o in oracles:
        o.reset()
    rs = np.random.RandomState(2)
    K, N = 4, cfg.num_envs
    from test_gpu_fork import push_state as push_fork_state
    for s in range(12):
        push_fork_state(sim, oracles, cfg)
        a = rs.uniform(-1, 1, (N * K, 2)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda())
        ref = []
        for e, o in enumerate(oracles):
            o.step(a[e * K:(e + 1) * K].astype(np.float64))
            ref.append(o.reward_info()[:, 0])
        got, ref = sim.reward_info_host()[:, 0], np.concatenate(ref)
        close = np.abs(got - ref) <= 2e-5 * np.abs(ref) + 2e-6        # a capture-radius tie ends the call one sub-step apart (rare)
        assert close.mean() >= 0.95 and (ref > 0).all(), (got[~close], ref[~close])


def test_facade_on_gpu_runs_an_episode():
    from quad_swarm_rl_stable_baselines3_b200.env import QuadrotorEnvMulti
    cfg = QuadSimConfig(num_envs=1, num_agents=8, ep_time=0.2, seed=2)
    env = QuadrotorEnvMulti(cfg, device="cuda:0")
    obs, info = env.reset()
    assert obs.shape == (8, 54) and info == {}
    rs = np.random.RandomState(0)
    n_done = 0
    for t in range(25):
        obs, rew, dones, infos = env.step(rs.uniform(-1, 1, (8, 4)))
        r = infos[3]["rewards"]
        total = r["rew_pos"] + r["rew_action"] + r["rew_crash"] + r["rew_orient"] + r["rew_spin"] + r["rew_quadcol"] + r["rew_proximity"]
        assert abs(total - rew[3]) < 1e-6
        n_done += int(dones[0])
        if dones[0]:
            assert "episode_extra_stats" in infos[0] and env.envs[0].tick == 0
    assert n_done == 1 and env.envs[2].dynamics.rot.shape == (3, 3)
    fcfg = QuadSimConfig.fork_default(num_envs=1, num_agents=4, ep_time=0.4, capture_radius=2.6, seed=3)
    fenv = QuadrotorEnvMulti(fcfg, device="cuda:0")
    obs, info = fenv.reset()
    assert info == {"success": False}
    for t in range(20):
        obs, rew, dones, infos = fenv.step(rs.uniform(-1, 1, (4, 2)))
        # quadrotor_single_rewards.py:457: every agent's info carries 'goal_dist' (what swarm_rl/sb_eval.py:28 averages) and an empty 'rewards'
        assert all(i["rewards"] == {} and 0.0 < i["goal_dist"] < 30.0 for i in infos)
        if not dones[0]:
            gd = [np.linalg.norm(e.dynamics.pos - e.goal) for e in fenv.envs]
            np.testing.assert_allclose([i["goal_dist"] for i in infos], gd, atol=0.02)    # the evader moves <= 5 mm per sub-step
        if dones[0]:
            obs2, info = fenv.reset()
            assert set(info) == {"success"} and not np.array_equal(obs, obs2)
            break
    else:
        raise AssertionError("no episode end in 20 calls")
