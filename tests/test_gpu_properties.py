"""Size-independent properties at the full BASELINE.json sizes, the non-finite-state guard, and the statistics of the
device noise (the oracle cannot be run at these sizes in seconds, so these complement the value-for-value parity tests)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402


def _sim(cfg):
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
    return QuadSwarmSim(cfg, device="cuda:0")


FULL = {
    # BASELINE.json configs[1..4] at their stated sizes (short episodes so that every env resets inside the run)
    "cfg2_4096x8": lambda: QuadSimConfig(num_envs=4096, num_agents=8, ep_time=0.25, seed=1),
    "cfg3_obst_4096x8": lambda: QuadSimConfig(num_envs=4096, num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                              obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, ep_time=0.25, seed=2),
    "cfg4_1024x32": lambda: QuadSimConfig(num_envs=1024, num_agents=32, ep_time=0.25, seed=3),
    "cfg5_65536x8": lambda: QuadSimConfig(num_envs=65536, num_agents=8, ep_time=0.25, seed=4),
    "fork_16384x4": lambda: QuadSimConfig.fork_default(num_envs=16384, num_agents=4, ep_time=0.8, capture_radius=2.6, seed=5),
}


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_invariants(name):
    import torch
    cfg = FULL[name]()
    sim = _sim(cfg)
    N, K, D = cfg.num_envs, cfg.num_agents, cfg.obs_dim
    obs = sim.reset()
    assert obs.shape == (N * K, D) and bool(torch.isfinite(obs).all())
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    hx, hy, hz = cfg.room_dims[0] / 2, cfg.room_dims[1] / 2, cfg.room_dims[2]
    ticks_before = sim.get_state(["tick"])["tick"].clone()
    per_call = cfg.fork.substeps if cfg.env_mode == "fork" else 1
    done_envs = 0
    for s in range(40):
        a = torch.rand((N * K, cfg.act_dim), device="cuda", generator=gen) * 2 - 1
        obs, rew, done = sim.step(a)
        st = sim.get_state(["pos", "rot", "tick", "omega"])
        d2 = done.view(N, K)
        assert bool((d2.all(dim=1) == d2.any(dim=1)).all()), "all agents of an env finish together"
        env_done = d2[:, 0]
        # done <=> the episode tick wrapped to 0; otherwise it advanced by the control steps of one call
        assert bool((st["tick"][env_done] == 0).all())
        if cfg.env_mode != "fork":
            assert bool((st["tick"][~env_done] == ticks_before[~env_done] + per_call).all())
            assert bool((env_done == (ticks_before + 1 > cfg.ep_len)).all())
        ticks_before = st["tick"].clone()
        done_envs += int(env_done.sum())
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
        p = st["pos"]
        assert bool((p[:, 0].abs() <= hx + 1e-6).all() and (p[:, 1].abs() <= hy + 1e-6).all())
        assert bool((p[:, 2] >= 0).all() and (p[:, 2] <= hz + 1e-6).all())
        # |omega| is clipped to 40 inside the dynamics; collision impulses add up to 20 pi per event after the clip
        assert bool((st["omega"].abs() <= 40.0 + 4 * 20 * np.pi).all())
        R = st["rot"].view(-1, 3, 3)
        err = (R @ R.transpose(1, 2) - torch.eye(3, device="cuda")).abs().amax()
        assert float(err) < 2e-4, f"rotation matrices drifted from orthonormal: {float(err)}"
        if cfg.env_mode != "fork":
            S = 18 if cfg.obs_repr == "xyz_vxyz_R_omega" else 19
            nb = obs[:, S:S + 6 * cfg.visible].view(N * K, cfg.visible, 6)
            assert bool((nb[..., 3:].abs() <= 6.0).all()) and bool((nb[..., 0].abs() <= cfg.room_dims[0]).all())
    assert done_envs >= N                                   # every env finished at least one episode
    stats = sim.episode_stats()
    assert stats["episodes"] == done_envs and stats["nonfinite_resets"] == 0


def test_two_shards_reproduce_the_full_run_bitwise():
    """env sharding at the cfg5 size: two handles of 32768 envs with env_id_offset 0 / 32768 == one handle of 65536."""
    import torch
    from quad_swarm_rl_stable_baselines3_b200.sharding import shard_config
    full_cfg = QuadSimConfig(num_envs=65536, num_agents=8, ep_time=0.1, seed=11)
    full = _sim(full_cfg)
    shards = [_sim(shard_config(full_cfg, r, 2)) for r in range(2)]
    o = full.reset()
    assert torch.equal(o, torch.cat([s.reset() for s in shards]))
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    half = 32768 * 8
    for t in range(14):
        a = torch.rand((65536 * 8, 4), device="cuda", generator=gen) * 2 - 1
        o, r, d = full.step(a)
        parts = [s.step(a[i * half:(i + 1) * half].contiguous()) for i, s in enumerate(shards)]
        assert torch.equal(o, torch.cat([p[0] for p in parts])) and torch.equal(r, torch.cat([p[1] for p in parts]))
        assert torch.equal(d, torch.cat([p[2] for p in parts]))
    tot = full.episode_stats()
    assert tot["episodes"] == sum(s.episode_stats()["episodes"] for s in shards) > 0


def test_nonfinite_state_forces_a_reset_and_is_counted():
    """A NaN in a drone's state cannot be stepped: that env is force-reset and counted (the reference raises ValueError,
    quadrotor_single.py:87-90); every other env is untouched (bitwise equal to a twin run)."""
    import torch
    cfg = QuadSimConfig(num_envs=64, num_agents=8, seed=5)
    a_sim, b_sim = _sim(cfg), _sim(cfg)
    a_sim.reset(); b_sim.reset()
    st = a_sim.get_state(["vel"])
    st["vel"][8 * 13 + 2, 1] = float("nan")                 # env 13, drone 2
    a_sim.set_state(vel=st["vel"])
    act = torch.rand((64 * 8, 4), device="cuda") * 2 - 1
    oa, ra, da = a_sim.step(act)
    ob, rb, db = b_sim.step(act)
    rows = torch.arange(64 * 8, device="cuda") // 8 == 13
    assert bool(da[rows].all()) and not bool(da[~rows].any())
    assert bool(torch.isfinite(oa).all())
    assert torch.equal(oa[~rows], ob[~rows]) and torch.equal(ra[~rows], rb[~rows])
    assert a_sim.episode_stats()["nonfinite_resets"] == 1 and b_sim.episode_stats()["nonfinite_resets"] == 0
    oa2, _, _ = a_sim.step(act)
    assert bool(torch.isfinite(oa2).all()) and bool(torch.isfinite(a_sim.get_state(["pos"])["pos"]).all())


def test_device_noise_statistics():
    """Philox + SFU Box-Muller: sensor noise N(0, 0.005 / 0.01 / 1.75e-4) on pos / vel / omega (sensor_noise.py:70-76) and
    the OU thrust noise (theta 0.15, sigma 0.01; numba_utils.py:77-105) have the reference's moments."""
    import torch
    cfg = QuadSimConfig(num_envs=8192, num_agents=8, seed=23)
    sim = _sim(cfg)
    sim.reset()
    n = 8192 * 8
    act = torch.zeros((n, 4), device="cuda")
    ou_prev = None
    for t in range(60):
        obs, _, _ = sim.step(act)
        if t == 58:
            ou_prev = sim.get_state(["ou"])["ou"].clone()
    st = sim.get_state(["pos", "vel", "omega", "goal", "ou"])
    npos = (obs[:, 0:3] + st["goal"] - st["pos"]).double()
    nvel = (obs[:, 3:6] - st["vel"]).double()
    nom = (obs[:, 15:18] - st["omega"]).double()
    for x, sigma in ((npos, 0.005), (nvel, 0.01), (nom, 0.000175)):
        m = x.numel()
        assert abs(float(x.mean())) < 5 * sigma / np.sqrt(m)
        assert abs(float(x.std()) / sigma - 1.0) < 0.01
        z = (x / sigma).flatten()
        assert abs(float((z ** 4).mean()) - 3.0) < 0.1                      # Gaussian kurtosis
        assert abs(float((z.abs() > 3).double().mean()) - 0.0027) < 0.0006  # tails
    ou = st["ou"].double()
    stat_std = 0.01 / np.sqrt(1 - (1 - 0.15) ** 2)
    assert abs(float(ou.std()) / stat_std - 1.0) < 0.02
    rho = float(((ou * ou_prev.double()).mean()) / (ou.std() * ou_prev.double().std()))
    assert abs(rho - 0.85) < 0.01                                           # x' = 0.85 x + sigma n
    c = np.corrcoef(npos[:, 0].cpu().numpy(), npos[:, 1].cpu().numpy())[0, 1]
    assert abs(c) < 0.02                                                    # independent components


@pytest.mark.parametrize("name,cfg_fn", [
    ("ragged_k3", lambda: QuadSimConfig(num_envs=37, num_agents=3, neighbor_visible_num=1, ep_time=0.05, seed=1)),
    ("ragged_k8", lambda: QuadSimConfig(num_envs=131, num_agents=8, ep_time=0.05, seed=2)),
    ("ragged_mix_k5", lambda: QuadSimConfig(num_envs=29, num_agents=5, quads_mode="mix", neighbor_visible_num=2, ep_time=0.05, seed=3)),
    ("ragged_obst_k6", lambda: QuadSimConfig(num_envs=21, num_agents=6, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                             obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, ep_time=0.05, seed=4)),
    ("ragged_k32", lambda: QuadSimConfig(num_envs=7, num_agents=32, ep_time=0.05, seed=5)),
    ("ragged_fork_k4", lambda: QuadSimConfig.fork_default(num_envs=45, num_agents=4, ep_time=0.3, capture_radius=2.6, seed=6)),
])
def test_output_buffers_are_never_overrun(name, cfg_fn):
    """compute-sanitizer is not available on the GPU pool, so the bounds of every caller-provided output are checked with guard
    bands: obs / rew / done / terminal_obs / reset_success / episode records sit in the middle of larger canary-filled buffers
    (batch sizes that are not multiples of the warp tile, lane groups that are not full), and after resets, steps and episode
    ends the canaries are intact.  Goes through the C-ABI directly."""
    import ctypes as C
    import torch
    from quad_swarm_rl_stable_baselines3_b200 import _capi
    cfg = cfg_fn()
    sim = _sim(cfg)
    L, h = _capi.lib(), sim._h
    N, K, D, A = cfg.num_envs, cfg.num_agents, cfg.obs_dim, cfg.act_dim
    n, PAD = N * K, 4096                                   # PAD elements of guard band on each side (16-byte multiples)

    def guarded(count, dtype, fill):
        buf = torch.full((count + 2 * PAD,), fill, dtype=dtype, device="cuda")
        return buf, buf[PAD:PAD + count]

    bufs = {k: guarded(c, dt, f) for k, (c, dt, f) in dict(
        obs=(n * D, torch.float32, -777.0), rew=(n, torch.float32, -777.0), done=(n, torch.uint8, 0xAB), term=(n * D, torch.float32, -777.0),
        succ=(N, torch.uint8, 0xAB), erec=(N * 20, torch.int32, -777), arec=(n * 4, torch.float32, -777.0)).items()}
    ptr = {k: C.c_void_p(v[1].data_ptr()) for k, v in bufs.items()}
    stream = sim._stream()
    assert L.qs_reset(h, None, ptr["obs"], stream) == 0
    g = torch.Generator(device="cuda").manual_seed(1)
    finished = 0
    for s in range(40):
        act = (torch.rand((n, A), device="cuda", generator=g) * 2 - 1).contiguous()
        assert L.qs_step(h, C.c_void_p(act.data_ptr()), ptr["obs"], ptr["rew"], ptr["done"], ptr["term"], ptr["succ"], stream) == 0
        assert L.qs_episode_records(h, ptr["erec"], ptr["arec"], stream) == 0
        finished += int(bufs["done"][1].any())
        if s == 20:                                        # a masked reset in the middle
            mask = (torch.arange(N, device="cuda") % 3 == 0).to(torch.uint8)
            assert L.qs_reset(h, C.c_void_p(mask.data_ptr()), ptr["obs"], stream) == 0
    torch.cuda.synchronize()
    assert finished >= 2
    for k, (buf, view) in bufs.items():
        fill = buf[0].item()
        assert bool((buf[:PAD] == fill).all()) and bool((buf[PAD + view.numel():] == fill).all()), f"{name}: guard band of {k} overwritten"
    assert bool(torch.isfinite(bufs["obs"][1]).all()) and bool((bufs["obs"][1] != -777.0).all())      # and the inside was written everywhere
    assert bool((bufs["rew"][1] != -777.0).all()) and bool((bufs["done"][1] <= 1).all())
