"""A minimal restatement of stable-baselines3's `VecEnv` abstract base and of the access patterns of
`OnPolicyAlgorithm.collect_rollouts` / `VecMonitor` / `evaluate_policy` -- TEST INFRASTRUCTURE ONLY.

stable-baselines3 is not installed in the build image (and there is no network), so the claim "sb_train.py's PPO consumes the
VecEnv unchanged" is exercised against this stand-in: the abstract-method set, `step()` composition, attribute names and the way
SB3 reads / WRITES `infos`, `reset_infos`, `_last_obs` and `terminal_observation` follow SB3 2.x
(`stable_baselines3/common/vec_env/base_vec_env.py`, `on_policy_algorithm.py:collect_rollouts`, `vec_monitor.py:step_wait`).
"""
import abc
import sys
import types

import numpy as np


class VecEnv(abc.ABC):
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]
        self.render_mode = None

    def _reset_seeds(self):
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self):
        self._options = [{} for _ in range(self.num_envs)]

    @abc.abstractmethod
    def reset(self): ...

    @abc.abstractmethod
    def step_async(self, actions): ...

    @abc.abstractmethod
    def step_wait(self): ...

    @abc.abstractmethod
    def close(self): ...

    @abc.abstractmethod
    def get_attr(self, attr_name, indices=None): ...

    @abc.abstractmethod
    def set_attr(self, attr_name, value, indices=None): ...

    @abc.abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...

    @abc.abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_images(self):
        raise NotImplementedError

    def render(self, mode=None):
        return None

    def seed(self, seed=None):
        return [None] * self.num_envs

    @property
    def unwrapped(self):
        return self

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


def install():
    """Register the stand-in as `stable_baselines3.common.vec_env.base_vec_env` (returns the names added to sys.modules)."""
    names = ["stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env",
             "stable_baselines3.common.vec_env.base_vec_env"]
    added = []
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
            added.append(n)
    sys.modules[names[-1]].VecEnv = VecEnv
    return added


def uninstall(added):
    for n in added:
        sys.modules.pop(n, None)


class RolloutBufferLike:
    """The arrays SB3's RolloutBuffer fills in `add` (buffers.py): copies of obs / actions / rewards / episode starts."""

    def __init__(self, n_steps, n_envs, obs_dim, act_dim):
        self.obs = np.zeros((n_steps, n_envs, obs_dim), np.float32)
        self.actions = np.zeros((n_steps, n_envs, act_dim), np.float32)
        self.rewards = np.zeros((n_steps, n_envs), np.float32)
        self.episode_starts = np.zeros((n_steps, n_envs), np.float32)
        self.pos = 0

    def add(self, obs, action, reward, episode_start):
        self.obs[self.pos] = np.array(obs)
        self.actions[self.pos] = np.array(action)
        self.rewards[self.pos] = np.array(reward)
        self.episode_starts[self.pos] = np.array(episode_start)
        self.pos += 1


def collect_rollouts(env, n_steps, policy, rng):
    """The env-facing part of OnPolicyAlgorithm.collect_rollouts: `_last_obs` is held across exactly one env.step, actions are
    clipped to the Box, `infos[idx]["terminal_observation"]` is read for every done row, a VecMonitor-style wrapper WRITES
    `infos[i]["episode"]` in place, and `reset_infos` is consumed by the curriculum callback."""
    last_obs = env.reset()
    episode_starts = np.ones((env.num_envs,), dtype=bool)
    buf = RolloutBufferLike(n_steps, env.num_envs, env.observation_space.shape[0], env.action_space.shape[0])
    ep_ret = np.zeros(env.num_envs)
    ep_len = np.zeros(env.num_envs, dtype=int)
    log = dict(episodes=0, terminal_obs=0, reset_infos=0, written=0)
    for _ in range(n_steps):
        actions = policy(last_obs, rng)
        clipped = np.clip(actions, env.action_space.low, env.action_space.high)
        new_obs, rewards, dones, infos = env.step(clipped)
        assert new_obs is not last_obs, "SB3 keeps _last_obs while stepping: the env must not hand back the same array"
        held = last_obs.copy()
        # VecMonitor.step_wait: per-row bookkeeping, writes into the info dict of finished rows
        ep_ret += rewards
        ep_len += 1
        for i in range(len(dones)):
            infos[i]["monitor_seen"] = i                     # in-place write: rows must not share a dict
            if dones[i]:
                infos[i] = dict(infos[i], episode={"r": float(ep_ret[i]), "l": int(ep_len[i])})
                ep_ret[i] = 0.0
                ep_len[i] = 0
                log["episodes"] += 1
        assert [infos[i]["monitor_seen"] for i in range(len(infos))] == list(range(len(infos))), "infos rows alias one dict"
        log["written"] += len(infos)
        for idx, done in enumerate(dones):
            if done and infos[idx].get("terminal_observation") is not None and not infos[idx].get("TimeLimit.truncated", False):
                assert infos[idx]["terminal_observation"].shape == env.observation_space.shape
                log["terminal_obs"] += 1
        for r in env.reset_infos:                            # CurriculumCallback._on_step (custom_callbacks.py:452-456)
            if r is not None:
                log["reset_infos"] += 1
        buf.add(last_obs, actions, rewards, episode_starts)
        assert np.array_equal(last_obs, held)
        last_obs = new_obs
        episode_starts = dones
    return buf, log
