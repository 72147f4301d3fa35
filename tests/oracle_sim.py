"""Test helper: the CPU oracle behind QuadSwarmSim's host interface, so the host-side logic (VecEnv facade, sharding,
stat reduction) can be exercised on a box without a GPU.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

from oracle import OracleEnv
from quad_swarm_rl_stable_baselines3_b200.config import PARAM_KEYS


class OracleSim:
    def __init__(self, cfg):
        self.cfg = cfg
        self.N, self.K, self.D, self.A = cfg.num_envs, cfg.num_agents, cfg.obs_dim, cfg.act_dim
        self.envs = [OracleEnv(cfg, i) for i in range(self.N)]      # gid = cfg.env_id_offset + i inside qo_create
        self.closed = False

    def reset_host(self):
        return np.concatenate([o.reset() for o in self.envs]).astype(np.float32)

    def step_host(self, actions, out=None, terminal_obs=None, reset_success=None):
        a = np.asarray(actions, dtype=np.float64).reshape(self.N, self.K, self.A)
        if out is None:
            out = (np.empty((self.N * self.K, self.D), np.float32), np.empty(self.N * self.K, np.float32),
                   np.empty(self.N * self.K, np.uint8))
        obs, rew, done = out
        K = self.K
        for e, o in enumerate(self.envs):
            r = o.step(a[e], want_terminal=True)
            sl = slice(e * K, (e + 1) * K)
            obs[sl], rew[sl], done[sl] = r[0], r[1], r[2]
            if r[2].any():
                if terminal_obs is not None:
                    terminal_obs[sl] = r[3]
                if reset_success is not None:
                    reset_success[e] = int(bool(o.last_reset_success))
        return obs, rew, done.view(np.bool_)

    def episode_records_host(self, env_out=None, agent_out=None):
        env = np.empty((self.N, 20), np.int32) if env_out is None else env_out
        agent = np.empty((self.N * self.K, 4), np.float32) if agent_out is None else agent_out
        for e, o in enumerate(self.envs):
            env[e], agent[e * self.K:(e + 1) * self.K] = o.record()
        return env, agent

    def set_capture_radius(self, v):
        for o in self.envs:
            o.set_param(PARAM_KEYS["capture_radius"], float(v))

    def set_rew_coeff(self, **kw):
        for o in self.envs:
            for k, v in kw.items():
                o.set_param(PARAM_KEYS[k], float(v))

    def episode_stats(self, reset=False):
        tot = {}
        for o in self.envs:
            for k, v in o.stats().items():
                tot[k] = tot.get(k, 0) + v
        return tot

    def close(self):
        self.closed = True

    # ---- what the QuadrotorEnvMulti facade needs beyond the VecEnv surface ---------------------------------------
    def get_state_host(self, fields=None):
        sts = [o.get_state() for o in self.envs]
        per_drone = ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "ou", "goal", "flags", "col_mask")
        out = {k: np.concatenate([np.asarray(s[k]).reshape(self.K, -1) for s in sts]).astype(
            np.int32 if k in ("flags", "col_mask") else np.float32) for k in per_drone}
        for k in ("flags", "col_mask"):
            out[k] = out[k].reshape(-1)
        for k in ("tick", "svd_ctr", "step_ctr"):
            out[k] = np.array([s[k] for s in sts], dtype=np.int32)
        return out if fields is None else {k: out[k] for k in fields}

    def set_state(self, **fields):
        K = self.K
        for e, o in enumerate(self.envs):
            o.set_state(**{k: np.asarray(v, dtype=np.float64).reshape(self.N * K, -1)[e * K:(e + 1) * K] for k, v in fields.items()})

    def enable_reward_info(self, on=True):
        self._ri = bool(on)
        return self._ri or None

    def reward_info_host(self):
        return np.concatenate([o.reward_info() for o in self.envs]).astype(np.float32)
