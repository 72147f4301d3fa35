"""SB3-style VecEnv facade over the device simulator: the drop-in for the reference's `SubprocVecEnvCustom`
(swarm_rl/env_wrappers/subproc_vec_env_custom.py:88-248) hosting `SB3QuadrotorEnv` (env_wrappers/sb3_quad_env.py:18-66).

Same surface and the same meaning per member:
  * `num_envs = n_envs * agents_per_env` rows, row = env * K + agent (:139, :145-147);
  * `reset() -> obs[N*K, D]`, `step_async(actions[N*K, A])`, `step_wait() -> (obs, rews[N*K], dones[N*K], infos)` with a
    flat list of N*K info dicts (:149-164); finished envs are already reset and every agent of such an env carries
    `info["terminal_observation"]` (:43-45);
  * `reset_infos`: one entry PER ENV, `None` if that env did not reset in the last call, else `{"success": bool}`
    (:41,46,52,152; read by CurriculumCallback, swarm_rl/custom_callbacks.py:452-456);
  * `batch` attribute (:113; custom_callbacks.py:445), `env_method / get_attr / set_attr / env_is_wrapped` indexed per env,
    not per agent (:226-237), `close()`.
Instead of one process and one pipe per env there is one `QuadSwarmSim` handle; SB3's numpy rollout loop is served by
`qs_step_host` through page-locked buffers (one H2D + one D2H per step); `step_tensor` / `reset_tensor` return CUDA tensors.

Buffers: observations alternate between two pinned arrays, so the array returned by step t stays valid until step t+2
(SB3 holds `_last_obs` across exactly one `env.step`).
"""
from __future__ import annotations

import math
from typing import Any, List, Optional, Sequence

import numpy as np

from .config import OBS_REPR_DIM, QS_ER_COUNT, QuadSimConfig, episode_extra_stats

try:                                             # the real base class when SB3 is installed (it is not in the build image)
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _SB3VecEnv
except Exception:                                # pragma: no cover - depends on the environment
    _SB3VecEnv = None

try:
    from gymnasium import spaces as _spaces
except Exception:                                # pragma: no cover
    _spaces = None


class _Box:
    """Minimal stand-in for gymnasium.spaces.Box (low / high / shape / dtype / sample / contains)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.shape}, {self.dtype})"


def _box(low, high):
    low, high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
    if _spaces is not None:
        return _spaces.Box(low, high, dtype=np.float32)
    return _Box(low, high)


def make_spaces(cfg: QuadSimConfig):
    """Per-agent observation / action `Box`es with the reference's bounds (quadrotor_single.py:278-349,
    quadrotor_single_rewards.py:317-411, quadrotor_control.py:37-49,74-86)."""
    L, W, H = (float(v) for v in cfg.room_dims)
    room = [L, W, H]
    vmax, wmax = 3.0, 40.0                      # dynamics.vxyz_max / omega_max (quadrotor_dynamics.py:48-49)
    comp = {
        "xyz": ([-v for v in room], room), "vxyz": ([-vmax] * 3, [vmax] * 3), "R": ([-1.0] * 9, [1.0] * 9),
        "omega": ([-wmax] * 3, [wmax] * 3), "floor": ([0.0], [H]), "wall": ([0.0] * 6, [5.0] * 6),
        "cdist": ([0.0], [L / 2]), "cdistdot": ([-vmax], [vmax]), "dist": ([-L / 2], [L / 2]), "distdot": ([-vmax], [vmax]),
        "angle": ([-math.pi], [math.pi]), "sangle": ([-1.0] * 2, [1.0] * 2), "angledot": ([-wmax], [wmax]),
        "ndist": ([-L / 2], [L / 2]), "nsangle": ([-1.0] * 2, [1.0] * 2),
        "aw": ([-math.pi], [math.pi]), "awdot": ([-wmax], [wmax]),
        "rxyz": ([-v for v in room], room), "rvxyz": ([-2 * vmax] * 3, [2 * vmax] * 3), "octmap": ([-10.0] * 9, [10.0] * 9),
    }
    names = cfg.obs_repr.split("_")
    nb = {"pos_vel": ["rxyz", "rvxyz"], "dist_angle": ["dist", "angle"], "dist_sangle": ["dist", "sangle"],
          "dist_angle_heading": ["dist", "angle", "angle"], "dist_sangle_sheading": ["dist", "sangle", "sangle"],
          "ndist_nsangle": ["dist", "sangle"], "none": []}            # quadrotor_single_rewards.py:368-393
    names = names + nb[cfg.neighbor_obs_type] * cfg.visible
    if cfg.use_obstacles:
        names.append("octmap")
    low = np.concatenate([comp[n][0] for n in names])
    high = np.concatenate([comp[n][1] for n in names])
    assert low.size == cfg.obs_dim, (low.size, cfg.obs_dim, OBS_REPR_DIM[cfg.obs_repr])
    A = cfg.act_dim
    return _box(low, high), _box(-np.ones(A), np.ones(A))


class _PlainVecEnv:
    """What the facade needs from SB3's VecEnv base when SB3 is absent: the attributes and the step() composition."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.render_mode = None
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        return [None] * self.num_envs

    def _reset_seeds(self):
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self):
        self._options = [{} for _ in range(self.num_envs)]

    @property
    def unwrapped(self):
        return self


_Base = _SB3VecEnv if _SB3VecEnv is not None else _PlainVecEnv


class AnnealSchedule:
    """`swarm_rl/env_wrappers/quad_utils.py:13-17`: a reward coefficient that grows linearly from 0 to `final_value` over
    `anneal_env_steps` training steps (the upstream recipe anneals quadcol_bin, quadcol_bin_smooth_max and quadcol_bin_obst,
    quad_utils.py:78-89)."""

    def __init__(self, coeff_name: str, final_value: float, anneal_env_steps: float):
        self.coeff_name, self.final_value, self.anneal_env_steps = coeff_name, float(final_value), float(anneal_env_steps)

    def value(self, approx_total_training_steps: float) -> float:
        return min(self.final_value * approx_total_training_steps / self.anneal_env_steps, self.final_value)   # reward_shaping.py:116


class QuadSwarmVecEnv(_Base):
    def __init__(self, cfg: QuadSimConfig, device=None, sim=None):
        """`sim`: an object with QuadSwarmSim's host interface (tests inject an oracle-backed one on CPU boxes)."""
        if sim is None:
            from .sim import QuadSwarmSim
            sim = QuadSwarmSim(cfg, device=device)
        self.sim = sim
        self.cfg = cfg
        self.batch = 0                                  # subproc_vec_env_custom.py:113
        self.waiting = False
        self.closed = False
        self.agents_per_env = cfg.num_agents
        self.n_envs = cfg.num_envs
        obs_space, act_space = make_spaces(cfg)
        super().__init__(cfg.num_envs * cfg.num_agents, obs_space, act_space)
        n, D = self.num_envs, cfg.obs_dim
        self._obs = [self._host_array((n, D), np.float32), self._host_array((n, D), np.float32)]
        self._flip = 0
        self._rew = self._host_array((n,), np.float32)
        self._done = self._host_array((n,), np.uint8)
        self._term = self._host_array((n, D), np.float32)
        self._succ = self._host_array((self.n_envs,), np.uint8)
        self._actions = self._host_array((n, cfg.act_dim), np.float32)
        self.reset_infos = tuple({} for _ in range(self.n_envs))
        self.capture_radius = float(cfg.fork.capture_radius)
        # per-episode records -> infos[i]['episode_extra_stats']; fetched only on steps where some env finished
        self.episode_infos = hasattr(sim, "episode_records_host")
        self._erec = np.zeros((self.n_envs, QS_ER_COUNT), dtype=np.int32)
        self._arec = np.zeros((n, 4), dtype=np.float32)

    @classmethod
    def from_reference_cfg(cls, rcfg, num_envs: int, device=None, **kw):
        """Build from a `swarm_rl.global_cfg.QuadrotorEnvConfig` the way sb_train.py does (sb_train.py:44-50)."""
        return cls(QuadSimConfig.from_reference_cfg(rcfg, num_envs, **kw), device=device)

    @staticmethod
    def _host_array(shape, dtype):
        try:                                            # page-locked so qs_step_host DMA's straight into it
            import torch
            if torch.cuda.is_available():
                tdt = {np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
                return torch.zeros(shape, dtype=tdt).pin_memory().numpy()
        except Exception:
            pass
        return np.zeros(shape, dtype=dtype)

    # ---- VecEnv interface -----------------------------------------------------------------------------------
    def reset(self):
        obs = self.sim.reset_host()
        self._flip ^= 1
        buf = self._obs[self._flip]
        buf[...] = obs
        self.reset_infos = tuple({"success": False} if self.cfg.env_mode == "fork" else {} for _ in range(self.n_envs))
        if hasattr(self, "_reset_seeds"):
            self._reset_seeds()
            self._reset_options()
        return buf

    def step_async(self, actions) -> None:
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != self._actions.shape:
            raise ValueError(f"actions must have shape {self._actions.shape}, got {a.shape}")
        self._actions[...] = a
        self.waiting = True

    def step_wait(self):
        self._flip ^= 1
        obs = self._obs[self._flip]
        self.sim.step_host(self._actions, (obs, self._rew, self._done), terminal_obs=self._term, reset_success=self._succ)
        self.waiting = False
        K = self.agents_per_env
        dones = self._done.view(np.bool_).copy()
        rews = self._rew.copy()
        env_done = dones.reshape(self.n_envs, K)[:, 0]
        infos: List[dict] = [{} for _ in range(self.num_envs)]      # one dict per agent row, as the reference's workers return
        reset_infos: List[Optional[dict]] = [None] * self.n_envs
        finished = np.flatnonzero(env_done)
        if finished.size and self.episode_infos:
            # infos[i]['episode_extra_stats'] of the episodes that just ended (quadrotor_multi.py:739-831)
            self.sim.episode_records_host(self._erec, self._arec)
        for e in finished:
            reset_infos[e] = {"success": bool(self._succ[e])} if self.cfg.env_mode == "fork" else {}
            for k in range(K):
                info = {"terminal_observation": self._term[e * K + k].copy()}
                if self.episode_infos:
                    info["episode_extra_stats"] = episode_extra_stats(self._erec[e], self._arec[e * K + k], K, self.cfg.use_obstacles)
                infos[e * K + k] = info
        self.reset_infos = tuple(reset_infos)
        return obs, rews, dones, infos

    def anneal_reward_coefficients(self, approx_total_training_steps: float, schedules: Sequence[AnnealSchedule]) -> dict:
        """The annealing step of `QuadsRewardShapingWrapper.step` (swarm_rl/env_wrappers/reward_shaping.py:109-118): set every
        scheduled coefficient to its value at this point of training (one `qs_set_param` each, effective from the next step).
        Returns the `z_anneal_<name>` entries the wrapper adds to `episode_extra_stats`."""
        values = {sch.coeff_name: sch.value(approx_total_training_steps) for sch in schedules}
        self.sim.set_rew_coeff(**values)
        return {f"z_anneal_{k}": v for k, v in values.items()}

    def close(self) -> None:
        if self.closed:
            return
        self.sim.close()
        self.closed = True

    def get_images(self) -> Sequence[Optional[np.ndarray]]:
        return [None for _ in range(self.n_envs)]               # rendering is out of scope

    def render(self, mode=None):
        return None

    def _get_indices(self, indices) -> List[int]:
        if indices is None:
            return list(range(self.n_envs))                     # per env, not per agent (:226-231)
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        idx = self._get_indices(indices)
        if method_name == "set_capture_radius":                 # sb3_quad_env.py:44-45, custom_callbacks.py:460-461
            value = float(method_args[0] if method_args else method_kwargs["value"])
            self.sim.set_capture_radius(value)                  # one radius for all envs of the handle
            self.capture_radius = value
            return [None for _ in idx]
        if method_name == "set_rew_coeff":
            self.sim.set_rew_coeff(**method_kwargs)
            return [None for _ in idx]
        raise AttributeError(f"env_method {method_name!r} is not available on the device simulator")

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = self._get_indices(indices)
        table = {"num_agents": self.agents_per_env, "capture_radius": self.capture_radius, "cfg": self.cfg,
                 "observation_space": self.observation_space, "action_space": self.action_space, "render_mode": None}
        if attr_name not in table:
            raise AttributeError(attr_name)
        return [table[attr_name] for _ in idx]

    def has_attr(self, attr_name: str) -> bool:
        try:
            self.get_attr(attr_name)
            return True
        except AttributeError:
            return False

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        if attr_name == "capture_radius":
            self.env_method("set_capture_radius", value, indices=indices)
            return
        raise AttributeError(f"set_attr {attr_name!r} is not available on the device simulator")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._get_indices(indices)]

    # ---- device-resident face (no host round trip) ----------------------------------------------------------
    def reset_tensor(self):
        return self.sim.reset()

    def step_tensor(self, actions):
        """actions: CUDA float32 [N*K, A] -> (obs, rew, done) CUDA tensors; `sim.terminal_obs` / `sim.reset_success` hold
        the finished envs' terminal observation / success flag."""
        return self.sim.step(actions)
