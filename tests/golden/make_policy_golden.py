#!/usr/bin/env python
"""Golden vectors for the policy encoder: the reference's own `QuadMultiEncoder`
(/root/reference/swarm_rl/models/quad_multi_model.py:250-354, neighbour encoders :23-140) is imported unmodified -- sample_factory's
three helpers it builds on are stubbed with their documented behaviour (fc_layer = nn.Linear, nonlinearity(cfg) = nn.Tanh for
cfg.nonlinearity == 'tanh', calc_num_elements = output size of a module) -- random weights are drawn, and its output on random
observations is recorded next to the weights.  tests/test_ppo_cpu.py loads the weights into `ppo.QuadEncoder` and must reproduce
the outputs.  Run here (the reference does not travel): python tests/golden/make_policy_golden.py"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def install_stubs():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    for n in ("sample_factory", "sample_factory.algo", "sample_factory.algo.utils", "sample_factory.model"):
        mod(n)
    ctx = mod("sample_factory.algo.utils.context")
    ctx.global_model_factory = lambda: types.SimpleNamespace(register_encoder_factory=lambda f: None)
    tu = mod("sample_factory.algo.utils.torch_utils")

    def calc_num_elements(module, shape):
        return int(module(torch.zeros((1,) + tuple(shape))).numel())
    tu.calc_num_elements = calc_num_elements
    enc = mod("sample_factory.model.encoder")

    class Encoder(nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.cfg = cfg
    enc.Encoder = Encoder
    mu = mod("sample_factory.model.model_utils")
    mu.fc_layer = lambda i, o, bias=True: nn.Linear(i, o, bias=bias)

    def nonlinearity(cfg, inplace=False):
        assert cfg.nonlinearity == "tanh"
        return nn.Tanh()
    mu.nonlinearity = nonlinearity
    # gym_art.quadrotor_multi.quad_utils imports numba / gymnasium at module scope in some revisions: only the three tables are needed
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import ref_harness as rh
    rh.install_stubs()


CASES = {
    # name: (obs_repr, neighbor_obs_type, num_agents, neighbor_visible_num, neighbor_encoder_type, use_obstacles)
    "mean_embed_k8": ("xyz_vxyz_R_omega", "pos_vel", 8, 6, "mean_embed", False),
    "attention_k8": ("xyz_vxyz_R_omega", "pos_vel", 8, 6, "attention", False),
    "mlp_k8": ("xyz_vxyz_R_omega", "pos_vel", 8, 6, "mlp", False),
    "mean_embed_obst": ("xyz_vxyz_R_omega_floor", "pos_vel", 8, 2, "mean_embed", True),
    "no_neighbours": ("xyz_vxyz_R_omega", "none", 1, 0, "mean_embed", False),
}


def main():
    install_stubs()
    from swarm_rl.models.quad_multi_model import QuadMultiEncoder
    out = {}
    torch.manual_seed(20261018)
    for name, (obs_repr, nbr, K, V, enc_type, obst) in CASES.items():
        cfg = types.SimpleNamespace(obs_repr=obs_repr, use_obstacles=obst, neighbor_hidden_size=48, neighbor_obs_type=nbr, num_agents=K,
                                    neighbor_visible_num=V, neighbor_encoder_type=enc_type, rnn_size=64, nonlinearity="tanh",
                                    obstacle_obs_type="octomap", obst_hidden_size=64)
        ref = QuadMultiEncoder(cfg, None)
        with torch.no_grad():
            for p in ref.parameters():
                p.copy_(torch.randn_like(p) * (0.4 / max(1.0, float(p.shape[-1]) ** 0.5) if p.dim() == 2 else 0.2))
        S = ref.self_obs_dim
        D = S + ref.all_neighbor_obs_size + (9 if obst else 0)
        obs = torch.randn(64, D) * 0.9
        with torch.no_grad():
            y = ref({"obs": obs})
        out[f"{name}/obs"] = obs.numpy()
        out[f"{name}/out"] = y.numpy()
        for k, v in ref.state_dict().items():
            out[f"{name}/w/{k}"] = v.numpy()
        print(name, "D", D, "out", tuple(y.shape), "params", sorted({k.split('.')[0] for k in ref.state_dict()}))
    path = os.path.join(HERE, "policy_encoder.npz")
    np.savez_compressed(path, **out)
    print("->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
