"""Host-side logic of the device PPO loop on CPU: GAE against a numpy restatement of SB3's
RolloutBuffer.compute_returns_and_advantage, policy shapes for every env family, and the two-rank (gloo) gradient
all-reduce -- driven through a CPU stand-in for the simulator (the oracle behind torch tensors)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig  # noqa: E402
from quad_swarm_rl_stable_baselines3_b200.ppo import DevicePPO, PPOConfig, QuadActorCritic, compute_gae  # noqa: E402


def test_gae_matches_sb3_recursion():
    rs = np.random.RandomState(0)
    T, n, gamma, lam = 17, 5, 0.99, 0.95
    rew, val = rs.randn(T, n), rs.randn(T, n)
    done = rs.rand(T, n) < 0.2
    last_v = rs.randn(n)
    # SB3: episode_starts[t] = dones[t-1]; next_non_terminal at step t = 1 - dones[t]
    adv = np.zeros((T, n))
    last = np.zeros(n)
    for t in reversed(range(T)):
        nnt = 1.0 - done[t]
        nv = last_v if t == T - 1 else val[t + 1]
        delta = rew[t] + gamma * nv * nnt - val[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    a, r = compute_gae(torch.tensor(rew), torch.tensor(val), torch.tensor(done), torch.tensor(last_v), gamma, lam)
    np.testing.assert_allclose(a.numpy(), adv, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r.numpy(), adv + val, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("cfg,enc", [
    (QuadSimConfig(num_envs=2, num_agents=8), "mean_embed"),
    (QuadSimConfig(num_envs=2, num_agents=8, quads_mode="mix", use_obstacles=True, obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2), "mlp"),
    (QuadSimConfig.fork_default(num_envs=2), "mean_embed"),
    (QuadSimConfig.fork_default(num_envs=2, num_agents=1, neighbor_obs_type="none", neighbor_visible_num=0), "mean_embed"),
])
def test_policy_shapes(cfg, enc):
    pol = QuadActorCritic(cfg, hidden=32, neighbor_hidden=16, neighbor_encoder=enc)
    obs = torch.randn(11, cfg.obs_dim)
    a, logp, v = pol.act(obs)
    assert a.shape == (11, cfg.act_dim) and logp.shape == (11,) and v.shape == (11,)
    lp, ent, v2 = pol.evaluate(obs, a)
    torch.testing.assert_close(lp, logp)
    assert float(pol.log_std.abs().max()) == 0.0          # log_std_init 0 (ActorCriticPolicyCustom.py:312)


class TensorOracleSim:
    """CPU stand-in with QuadSwarmSim's tensor interface."""

    def __init__(self, cfg):
        sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from oracle_sim import OracleSim
        self.o = OracleSim(cfg)
        self.device = torch.device("cpu")

    def reset(self):
        return torch.from_numpy(self.o.reset_host())

    def step(self, a):
        obs, rew, done = self.o.step_host(a.numpy())
        return torch.from_numpy(obs.copy()), torch.from_numpy(rew.copy()), torch.from_numpy(done.copy())

    def episode_stats(self, reset=False, reduce=False):
        st = self.o.episode_stats()
        if reduce:
            from quad_swarm_rl_stable_baselines3_b200.sharding import all_reduce_stats
            st = all_reduce_stats(st)
        return st


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from quad_swarm_rl_stable_baselines3_b200.sharding import shard_config
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = QuadSimConfig.fork_default(num_envs=4, num_agents=2, ep_time=0.4, seed=3)
    cfg = shard_config(full, rank, world)
    ppo = DevicePPO(TensorOracleSim(cfg), cfg, PPOConfig(n_steps=6, batch_size=16, n_epochs=2, hidden=16, neighbor_hidden=8), seed=1)
    w0 = torch.cat([q.detach().reshape(-1).clone() for q in ppo.policy.parameters()])
    torch.manual_seed(100 + rank)                              # different action noise / minibatch order per rank
    hist = ppo.learn(2)
    w1 = torch.cat([q.detach().reshape(-1) for q in ppo.policy.parameters()])
    torch.save(dict(w0=w0, w1=w1, hist=hist), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_keeps_replicas_identical(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(os.path.join(str(tmp_path), f"r{r}.pt"), weights_only=False) for r in range(2))
    assert torch.equal(a["w0"], b["w0"])                       # same seed -> same initial weights
    assert not torch.equal(a["w0"], a["w1"])                   # the update moved them
    torch.testing.assert_close(a["w1"], b["w1"], rtol=0, atol=1e-6)   # averaged gradients -> replicas stay in lock-step
    assert a["hist"][-1]["agent_steps"] == 2 * 2 * (2 * 2) * 6 and all(np.isfinite(r["pg"]) for r in a["hist"])
    assert a["hist"][-1]["episodes"] == b["hist"][-1]["episodes"]      # episode stats were reduced over both ranks


# ---- the encoder against the reference's own module ---------------------------------------------------------------------
# tests/golden/policy_encoder.npz (tests/golden/make_policy_golden.py): weights, observations and outputs of the unmodified
# swarm_rl/models/quad_multi_model.py:QuadMultiEncoder for every neighbour encoder type.  Parameter-name map reference -> ours.
_REF_TO_OURS = {
    "self_encoder.": "self_encoder.", "feed_forward.": "feed_forward.", "obstacle_encoder.": "obstacle.",
    "neighbor_encoder.embedding_mlp.": "neighbor.", "neighbor_encoder.neighbor_mlp.": "neighbor.",
    "neighbor_encoder.neighbor_value_mlp.": "neighbor_value.", "neighbor_encoder.attention_mlp.": "attention.",
}
_ENCODER_CASES = {
    "mean_embed_k8": (dict(num_agents=8), "mean_embed"),
    "attention_k8": (dict(num_agents=8), "attention"),
    "mlp_k8": (dict(num_agents=8), "mlp"),
    "mean_embed_obst": (dict(num_agents=8, quads_mode="mix", use_obstacles=True, obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2), "mean_embed"),
    "no_neighbours": (dict(num_agents=1, neighbor_obs_type="none", neighbor_visible_num=0), "mean_embed"),
}


@pytest.mark.parametrize("name", list(_ENCODER_CASES))
def test_encoder_matches_the_reference_module(name, golden_dir):
    import os
    from quad_swarm_rl_stable_baselines3_b200.ppo import QuadEncoder
    g = np.load(os.path.join(golden_dir, "policy_encoder.npz"))
    kw, kind = _ENCODER_CASES[name]
    enc = QuadEncoder(QuadSimConfig(num_envs=2, **kw), hidden=64, neighbor_hidden=48, neighbor_encoder=kind)
    sd = {}
    for key in g.files:
        if not key.startswith(name + "/w/"):
            continue
        ref = key[len(name) + 3:]
        pre = next(p for p in _REF_TO_OURS if ref.startswith(p))
        sd[_REF_TO_OURS[pre] + ref[len(pre):]] = torch.from_numpy(g[key])
    enc.load_state_dict(sd, strict=True)                          # every parameter of ours has a counterpart in the reference, and vice versa
    with torch.no_grad():
        y = enc(torch.from_numpy(g[name + "/obs"]))
    np.testing.assert_allclose(y.numpy(), g[name + "/out"], rtol=0, atol=2e-6)
