"""`QuadrotorEnvMulti`-shaped drop-in over the device simulator ("Env" row).

The reference's environment object is what eval / render scripts and `SB3QuadrotorEnv` hold
(swarm_rl/env_wrappers/sb3_quad_env.py:18-66, swarm_rl/raw_test.py:25-60): one swarm of K drones with

    reset()            -> (obs[K, D], info)          fork: info = {"success": bool} (quadrotor_multi_rewards.py:541,625-627)
                                                      upstream: info = {}           (quadrotor_multi.py:440,516)
    step(actions[K,A]) -> (obs[K, D], rewards: list[K], dones: list[K] of bool, infos: list[K] of dict)
                                                      (quadrotor_multi_rewards.py:632,992; quadrotor_multi.py:521,839)
    num_agents, observation_space, action_space (per-agent Box), envs[i].dynamics.{pos,vel,rot,omega}, envs[i].goal,
    set_capture_radius(v) (quadrotor_multi_rewards.py:210-211), all_dynamics(), close(), render()

`QuadrotorEnvMulti` here is that object on top of ONE simulator handle with `num_envs=1`; everything numeric happens in
csrc/quadsim_kernels.cuh.  Episode-boundary semantics follow the reference exactly:

  * upstream env: `step` resets the env itself when the episode ends and returns the first observation of the NEW episode
    with dones all True (quadrotor_multi.py:836-838) -- the kernel's in-step auto-reset is that reset;
  * fork env: `step` returns the LAST observation of the finished episode and does not reset; the caller (the VecEnv worker,
    subproc_vec_env_custom.py:43-46) then calls `reset()`, which reports `{"success": ...}` of the episode that just ended.
    The kernel has already performed that reset inside the step, so `step` hands out the terminal observation and the
    following `reset()` hands out the stashed first observation instead of launching a second reset.

This facade is a convenience for single-swarm use (evaluation, debugging, API compatibility); training goes through
`QuadSwarmVecEnv` / the tensor interface, which batch thousands of these per launch.
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from .config import ER, QS_ER_COUNT, QuadSimConfig, episode_extra_stats
from .vec_env import make_spaces

# the raw reward terms the kernel reports per drone-step (qs_set_reward_info): order of the 8-float row
REW_INFO_FIELDS = ("rewraw_pos", "rewraw_action", "rewraw_crash", "rewraw_orient", "rewraw_spin", "rewraw_quadcol",
                   "rew_proximity", "rewraw_quadcol_obstacle")


def reward_info_dict(row, rew_coeff: dict, use_obstacles: bool) -> dict:
    """infos[i]["rewards"] of the upstream env (quadrotor_single.py:69-84 scaled by dt, quadrotor_multi.py:642-649) from the
    kernel's raw terms.  `row` = the 8 floats of REW_INFO_FIELDS (raw terms are already multiplied by dt like the reference's)."""
    r = {k: float(v) for k, v in zip(REW_INFO_FIELDS, row)}
    out = {
        "rew_main": rew_coeff["pos"] * r["rewraw_pos"], "rew_pos": rew_coeff["pos"] * r["rewraw_pos"],
        "rew_action": rew_coeff["effort"] * r["rewraw_action"], "rew_crash": rew_coeff["crash"] * r["rewraw_crash"],
        "rew_orient": rew_coeff["orient"] * r["rewraw_orient"], "rew_spin": rew_coeff["spin"] * r["rewraw_spin"],
        "rewraw_main": r["rewraw_pos"], "rewraw_pos": r["rewraw_pos"], "rewraw_action": r["rewraw_action"],
        "rewraw_crash": r["rewraw_crash"], "rewraw_orient": r["rewraw_orient"], "rewraw_spin": r["rewraw_spin"],
        "rew_quadcol": rew_coeff["quadcol_bin"] * r["rewraw_quadcol"], "rew_proximity": r["rew_proximity"],
        "rewraw_quadcol": r["rewraw_quadcol"],
    }
    if use_obstacles:
        out["rew_quadcol_obstacle"] = rew_coeff["quadcol_bin_obst"] * r["rewraw_quadcol_obstacle"]
        out["rewraw_quadcol_obstacle"] = r["rewraw_quadcol_obstacle"]
    return out


class _DynamicsView:
    """`env.envs[i].dynamics` of the reference (QuadrotorDynamics attributes read by eval scripts, raw_test.py:39-42):
    live views of the device state of drone i, fetched through `qs_get_state` on access."""

    def __init__(self, owner: "QuadrotorEnvMulti", i: int):
        self._o, self._i = owner, i

    @property
    def pos(self) -> np.ndarray:
        return self._o._state()["pos"][self._i].astype(np.float64)

    @property
    def vel(self) -> np.ndarray:
        return self._o._state()["vel"][self._i].astype(np.float64)

    @property
    def rot(self) -> np.ndarray:
        return self._o._state()["rot"][self._i].astype(np.float64).reshape(3, 3)

    @property
    def omega(self) -> np.ndarray:
        return self._o._state()["omega"][self._i].astype(np.float64)

    @property
    def on_floor(self) -> bool:
        return bool(int(self._o._state()["flags"][self._i]) & 1)

    def set_state(self, position, velocity, rotation, omega, thrusts=None):
        """QuadrotorDynamics.set_state (quadrotor_dynamics.py:180-192) for this drone."""
        st = self._o._state()
        pos, vel, rot, om = st["pos"].copy(), st["vel"].copy(), st["rot"].copy(), st["omega"].copy()
        pos[self._i], vel[self._i] = np.asarray(position, np.float32), np.asarray(velocity, np.float32)
        rot[self._i], om[self._i] = np.asarray(rotation, np.float32).reshape(9), np.asarray(omega, np.float32)
        self._o.sim.set_state(pos=pos, vel=vel, rot=rot, omega=om)
        self._o._invalidate()


class _DroneView:
    """`env.envs[i]` (QuadrotorSingle): the members callers of the multi-env read."""

    def __init__(self, owner: "QuadrotorEnvMulti", i: int):
        self._o, self._i = owner, i
        self.dynamics = _DynamicsView(owner, i)

    @property
    def goal(self) -> np.ndarray:
        return self._o._state()["goal"][self._i].astype(np.float64)

    @property
    def tick(self) -> int:
        return int(self._o._state()["tick"][0])


class QuadrotorEnvMulti:
    def __init__(self, cfg=None, device=None, sim=None, **kwargs):
        """cfg: a `QuadSimConfig`, or the fork's `swarm_rl.global_cfg.QuadrotorEnvConfig` (duck-typed: as
        `QuadrotorEnvMulti(cfg=cfg)` in sb3_quad_env.py:39), or None with the upstream constructor's keyword arguments
        (`num_agents`, `ep_time`, `obs_repr`, `neighbor_visible_num`, ..., quadrotor_multi.py:27-45); arguments that only concern
        rendering / replay / numba (`use_numba`, `quads_render`, `quads_view_mode`, `use_replay_buffer`, `render_mode`, ...) are
        accepted and ignored.  `sim`: an object with QuadSwarmSim's host interface (tests inject an oracle-backed one)."""
        if isinstance(cfg, QuadSimConfig):
            scfg = cfg
        elif cfg is not None and hasattr(cfg, "neighbor_obs_type") and hasattr(cfg, "episode_duration"):
            scfg = QuadSimConfig.from_reference_cfg(cfg, 1)
        else:
            known = {f for f in QuadSimConfig.__dataclass_fields__}
            mapped = {k: v for k, v in kwargs.items() if k in known}
            if "rew_coeff" in mapped and mapped["rew_coeff"] is None:
                mapped.pop("rew_coeff")
            if cfg is not None and getattr(cfg, "seed", None) is not None:
                mapped.setdefault("seed", int(cfg.seed))
            scfg = QuadSimConfig(**mapped)
        if scfg.num_envs != 1:
            import dataclasses
            scfg = dataclasses.replace(scfg, num_envs=1)
        self.cfg = scfg
        if sim is None:
            from .sim import QuadSwarmSim
            sim = QuadSwarmSim(scfg, device=device)
        self.sim = sim
        self.num_agents = scfg.num_agents
        self.is_fork = scfg.env_mode == "fork"
        self.observation_space, self.action_space = make_spaces(scfg)
        self.envs = [_DroneView(self, i) for i in range(self.num_agents)]
        self.capture_radius = float(scfg.fork.capture_radius)
        self.episode_success = False
        K, D = self.num_agents, scfg.obs_dim
        self._obs = np.zeros((K, D), np.float32)
        self._rew = np.zeros((K,), np.float32)
        self._done = np.zeros((K,), np.uint8)
        self._term = np.zeros((K, D), np.float32)
        self._succ = np.zeros((1,), np.uint8)
        self._erec = np.zeros((1, QS_ER_COUNT), np.int32)
        self._arec = np.zeros((K, 4), np.float32)
        self._pending: Optional[np.ndarray] = None      # fork: first observation of the episode the kernel already started
        self._cache = None
        self._rew_info = None
        if hasattr(sim, "enable_reward_info"):                       # fork mode: column 0 carries goal_dist (QS_RI_RAW_POS)
            self._rew_info = sim.enable_reward_info()

    # ---- state views -----------------------------------------------------------------------------------------------
    def _state(self):
        if self._cache is None:
            st = self.sim.get_state_host(("pos", "vel", "rot", "omega", "goal", "flags", "tick"))
            self._cache = st
        return self._cache

    def _invalidate(self):
        self._cache = None

    def all_dynamics(self):
        return tuple(e.dynamics for e in self.envs)

    def set_capture_radius(self, value):
        self.capture_radius = float(value)
        self.sim.set_capture_radius(float(value))

    # ---- reset / step ----------------------------------------------------------------------------------------------
    def reset(self, obst_density=None, obst_size=None):
        self._invalidate()
        if self.is_fork:
            info = {"success": bool(self.episode_success)}
            self.episode_success = False
            if self._pending is not None:
                obs, self._pending = self._pending, None
                return obs.astype(np.float64), info
            obs = np.asarray(self.sim.reset_host(), np.float32).reshape(self.num_agents, -1)
            return obs.astype(np.float64), info
        obs = np.asarray(self.sim.reset_host(), np.float32).reshape(self.num_agents, -1)
        return obs.astype(np.float64), {}

    def step(self, actions):
        K = self.num_agents
        a = np.atleast_2d(np.asarray(actions, dtype=np.float32))
        if a.shape != (K, self.cfg.act_dim):
            raise ValueError(f"actions must have shape {(K, self.cfg.act_dim)}, got {a.shape}")
        self._invalidate()
        if self._pending is not None:
            self._pending = None          # the caller stepped on without reset(): the new episode is simply under way
            self.episode_success = False
        self.sim.step_host(a, (self._obs, self._rew, self._done), terminal_obs=self._term, reset_success=self._succ)
        done = bool(self._done[0])
        infos: List[dict] = [{} for _ in range(K)]
        if self._rew_info is not None:
            rows = self.sim.reward_info_host()
            if self.is_fork:
                # the fork's step returns {'rewards': dict(), 'goal_dist': |pos - goal|} per agent (quadrotor_single_rewards.py:457),
                # read by swarm_rl/sb_eval.py:28; the kernel evaluates it before the worker's auto-reset, as the reference does
                for i in range(K):
                    infos[i]["rewards"] = {}
                    infos[i]["goal_dist"] = float(rows[i][0])
            else:
                rc = self._rew_coeff()
                for i in range(K):
                    infos[i]["rewards"] = reward_info_dict(rows[i], rc, self.cfg.use_obstacles)
        if done and hasattr(self.sim, "episode_records_host"):
            self.sim.episode_records_host(self._erec, self._arec)
            for i in range(K):
                infos[i]["episode_extra_stats"] = episode_extra_stats(self._erec[0], self._arec[i], K, self.cfg.use_obstacles)
        rewards = [float(x) for x in self._rew]
        dones = [done] * K
        if self.is_fork and done:
            self.episode_success = bool(self._succ[0])
            self._pending = self._obs.copy()
            return self._term.astype(np.float64), rewards, dones, infos
        return self._obs.astype(np.float64), rewards, dones, infos

    def _rew_coeff(self):
        from .config import DEFAULT_REW_COEFF
        rc = dict(DEFAULT_REW_COEFF)
        rc.update({k: float(v) for k, v in self.cfg.rew_coeff.items() if k in rc})
        rc.update(getattr(self.sim, "rew_coeff_overrides", {}))
        return rc

    def render(self, *a, **k):
        return None                                   # rendering is out of scope

    def close(self):
        self.sim.close()


QuadrotorEnvMultiB200 = QuadrotorEnvMulti


class SB3QuadrotorEnv:
    """swarm_rl/env_wrappers/sb3_quad_env.py:18-66 on the device simulator: gymnasium-style 5-tuple step, `reset(seed, options)`."""

    def __init__(self, cfg=None, device=None, sim=None, **kwargs):
        self.cfg = cfg
        self.env = QuadrotorEnvMulti(cfg, device=device, sim=sim, **kwargs)
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space

    def set_capture_radius(self, value):
        self.env.set_capture_radius(value)

    def reset(self, seed=None, options=None):
        return self.env.reset()

    def step(self, action):
        obs, reward, terminated, info = self.env.step(action)
        return obs, reward, terminated, terminated, info

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()
