"""ctypes wrapper of the CPU oracle (oracle/quadsim_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from quad_swarm_rl_stable_baselines3_b200.config import QsConfigC, QsStatsC, QuadSimConfig, cube_floor_dim  # noqa: E402

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libquadsim_oracle.so")
    src = os.path.join(_HERE, "quadsim_oracle.c")
    hdr = os.path.join(_ROOT, "include", "quadsim.h")
    if os.path.exists(src):      # make tracks the .c / .inc / header dependencies; without sources keep the prebuilt library
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []) + ["libquadsim_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.qo_create.restype = C.c_void_p
        L.qo_create.argtypes = [C.POINTER(QsConfigC), C.c_int]
        L.qo_destroy.argtypes = [C.c_void_p]
        L.qo_obs_dim.argtypes = [C.c_void_p]
        L.qo_set_tape.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, dp, C.c_int]
        L.qo_tape_pos.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.qo_reset.argtypes = [C.c_void_p, dp]
        L.qo_step.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_uint8), dp, C.POINTER(C.c_uint8)]
        L.qo_act_dim.argtypes = [C.c_void_p]
        L.qo_get_fork_state.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int32)]
        L.qo_set_fork_state.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int32)]
        L.qo_dynamics_only.argtypes = [C.c_void_p, C.c_int, dp]
        L.qo_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        L.qo_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        L.qo_get_scenario.argtypes = [C.c_void_p, dp]
        L.qo_set_scenario.argtypes = [C.c_void_p, dp]
        L.qo_generate_goals.argtypes = [C.c_int, C.c_double, C.c_int, dp, C.c_double, C.c_int, dp]
        L.qo_set_obstacles.argtypes = [C.c_void_p, dp, C.c_int]
        L.qo_get_obstacles.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int)]
        L.qo_get_stats.argtypes = [C.c_void_p, C.POINTER(QsStatsC)]
        L.qo_col_norm_and_new_vel_obst.argtypes = [dp, dp, dp, dp]
        L.qo_col_norm_and_new_vel_obst.restype = C.c_double
        L.qo_get_record.argtypes = [C.c_void_p, C.c_void_p, dp]
        L.qo_get_diag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qo_set_param.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.qo_get_margins.argtypes = [C.c_void_p, dp]
        L.qo_get_reward_info.argtypes = [C.c_void_p, dp]
        L.qo_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.qo_batch_step.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp, dp, dp, C.POINTER(C.c_uint8)]
        L.qo_batch_reset.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp]
        L.qo_batch_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint8), C.c_int]
        L.qo_batch_run.restype = C.c_longlong
        L.qo_omp_threads.restype = C.c_int
        L.qo_set_tick.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def philox(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().qo_philox(c0, c1, c2, c3, k0, k1, out)
    return np.array(list(out), dtype=np.uint32)


STATE_FIELDS = (("pos", 3, np.float64), ("vel", 3, np.float64), ("rot", 9, np.float64), ("omega", 3, np.float64),
                ("rot_damp", 4, np.float64), ("cmds_damp", 4, np.float64), ("ou", 4, np.float64),
                ("goal", 3, np.float64), ("flags", 1, np.int32), ("col_mask", 1, np.uint32))


class OracleEnv:
    """One K-drone environment, float64, reference iteration order."""

    def __init__(self, cfg: QuadSimConfig, env_index: int = 0):
        self.cfg = cfg
        self.c = cfg.to_c()
        self.K = cfg.num_agents
        self.h = lib().qo_create(C.byref(self.c), env_index)
        if not self.h:
            raise RuntimeError("qo_create failed")
        self.D = lib().qo_obs_dim(self.h)
        self.A = lib().qo_act_dim(self.h)
        self.last_reset_success = None
        self._tape = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_destroy(self.h)
            self.h = None

    def set_tape(self, normals=None, uniforms=None, choices=None):
        """Replay recorded unit draws (reference order).  Pass nothing to go back to Philox."""
        if normals is None and uniforms is None and choices is None:
            self._tape = None
            lib().qo_set_tape(self.h, None, 0, None, 0, None, 0)
            return
        n = np.ascontiguousarray(normals if normals is not None else [], dtype=np.float64)
        u = np.ascontiguousarray(uniforms if uniforms is not None else [], dtype=np.float64)
        c = np.ascontiguousarray(choices if choices is not None else [], dtype=np.float64)
        self._tape = (n, u, c)          # keep alive
        # non-NULL pointers even when empty so that tape mode stays on
        lib().qo_set_tape(self.h, _dp(n) if n.size else C.cast(C.c_void_p(8), C.POINTER(C.c_double)), n.size,
                          _dp(u) if u.size else C.cast(C.c_void_p(8), C.POINTER(C.c_double)), u.size,
                          _dp(c) if c.size else C.cast(C.c_void_p(8), C.POINTER(C.c_double)), c.size)

    def tape_pos(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().qo_tape_pos(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def reset(self):
        obs = np.zeros((self.K, self.D))
        lib().qo_reset(self.h, _dp(obs))
        return obs

    def step(self, actions, want_terminal=False):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.K, self.A)
        obs = np.zeros((self.K, self.D))
        rew = np.zeros(self.K)
        done = np.zeros(self.K, dtype=np.uint8)
        term = np.zeros((self.K, self.D)) if want_terminal else None
        succ = C.c_uint8(255)
        lib().qo_step(self.h, _dp(a), _dp(obs), _dp(rew), done.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(term), C.byref(succ))
        self.last_reset_success = None if succ.value == 255 else bool(succ.value)   # reset_info["success"] if the env reset
        if want_terminal:
            return obs, rew, done.astype(bool), term
        return obs, rew, done.astype(bool)

    def dynamics_only(self, drone, thrust_cmd01):
        a = np.ascontiguousarray(thrust_cmd01, dtype=np.float64)
        lib().qo_dynamics_only(self.h, drone, _dp(a))

    def get_state(self):
        out = {n: np.zeros((self.K, w) if w > 1 else (self.K,), dtype=dt) for n, w, dt in STATE_FIELDS}
        tss = np.zeros(3, dtype=np.int32)
        lib().qo_get_state(self.h, *[_vp(out[n]) for n, _, _ in STATE_FIELDS], _vp(tss))
        out["tick"], out["svd_ctr"], out["step_ctr"] = int(tss[0]), int(tss[1]), int(tss[2])
        xy = np.zeros((64, 2))
        n = C.c_int()
        lib().qo_get_obstacles(self.h, _dp(xy), C.byref(n))
        out["obst_xy"] = xy[:n.value].copy()
        return out

    def set_state(self, **kw):
        ptrs, keep = [], []
        for n, w, dt in STATE_FIELDS:
            if n in kw and kw[n] is not None:
                a = np.ascontiguousarray(kw[n], dtype=dt).reshape((self.K, w) if w > 1 else (self.K,))
                keep.append(a)
                ptrs.append(_vp(a))
            else:
                ptrs.append(None)
        tss = None
        if any(k in kw for k in ("tick", "svd_ctr", "step_ctr")):
            cur = self.get_state()
            tss = np.array([kw.get("tick", cur["tick"]), kw.get("svd_ctr", cur["svd_ctr"]),
                            kw.get("step_ctr", cur["step_ctr"])], dtype=np.int32)
        lib().qo_set_state(self.h, *ptrs, _vp(tss))
        if kw.get("obst_xy") is not None:
            xy = np.ascontiguousarray(kw["obst_xy"], dtype=np.float64).reshape(-1, 2)
            lib().qo_set_obstacles(self.h, _dp(xy), xy.shape[0])

    def get_scenario(self):
        """QS_SC_* row (include/quadsim.h) of the formation-scenario state."""
        o = np.zeros(24)
        lib().qo_get_scenario(self.h, _dp(o))
        return o

    def set_scenario(self, row):
        a = np.ascontiguousarray(row, dtype=np.float64).reshape(24)
        lib().qo_set_scenario(self.h, _dp(a))

    def get_fork_state(self):
        pid, heading, evader, fl = np.zeros((self.K, 24)), np.zeros((self.K, 3)), np.zeros(2), (C.c_int32 * 2)()
        lib().qo_get_fork_state(self.h, _dp(pid), _dp(heading), _dp(evader), fl)
        return dict(pid=pid, heading=heading, evader=evader, episode_success=bool(fl[0]), chasers_placed=bool(fl[1]))

    def set_fork_state(self, pid=None, heading=None, evader=None, episode_success=None, chasers_placed=None):
        keep = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (pid, heading, evader)]
        fl = None
        if episode_success is not None or chasers_placed is not None:
            cur = self.get_fork_state()
            fl = (C.c_int32 * 2)(int(cur["episode_success"] if episode_success is None else episode_success),
                                 int(cur["chasers_placed"] if chasers_placed is None else chasers_placed))
        lib().qo_set_fork_state(self.h, *[_dp(a) for a in keep], fl)

    def stats(self):
        s = QsStatsC()
        lib().qo_get_stats(self.h, C.byref(s))
        return s.as_dict()

    def record(self):
        """(env_rec int32 [QS_ER_COUNT], agent_rec float64 [K, 4]) of the last finished episode (include/quadsim.h QS_ER_*)."""
        env_rec, agent_rec = np.zeros(20, dtype=np.int32), np.zeros((self.K, 4))
        lib().qo_get_record(self.h, _vp(env_rec), _dp(agent_rec))
        return env_rec, agent_rec

    def diag(self):
        new_pairs = np.zeros(self.K, dtype=np.uint32)
        V = max(self.cfg.visible, 1)
        nb = np.zeros((self.K, V), dtype=np.int32)
        flag = C.c_int32()
        lib().qo_get_diag(self.h, _vp(new_pairs), _vp(nb), C.byref(flag))
        return dict(new_pairs=new_pairs, neighbors=nb[:, :self.cfg.visible], impulse_flag=int(flag.value))

    def reward_info(self):
        """Raw reward terms of the last step, [K, 8] in QS_RI_* order (infos[i]["rewards"]; fork mode: goal_dist in column 0)."""
        out = np.zeros((self.K, 8))
        lib().qo_get_reward_info(self.h, _dp(out))
        return out

    MARGIN_CLASSES = ("pair", "falloff", "obst", "floor", "wall", "ceil", "yaw", "rank", "reach", "liftoff")

    def margins(self):
        """Distance of the last step's threshold decisions from their thresholds, per class, in fp32 ulps of the operands'
        magnitude (min over the env; 1e300 = the class made no decision).  A discrete disagreement with the fp32 kernel is a
        tie only if the class that produces that flag is within a few ulps here."""
        m = np.zeros(lib().qo_margin_count())
        lib().qo_get_margins(self.h, _dp(m))
        return dict(zip(self.MARGIN_CLASSES, m))

    def set_param(self, key: int, value: float):
        lib().qo_set_param(self.h, key, value)


def col_norm_and_new_vel_obst(pos, vel, obst_pos):
    """compute_col_norm_and_new_vel_obst of the oracle (collisions/obstacles.py:8-21) -> (vnew, collision_norm)."""
    p, v, o = (np.ascontiguousarray(x, dtype=np.float64) for x in (pos, vel, obst_pos))
    n = np.zeros(3)
    vnew = lib().qo_col_norm_and_new_vel_obst(_dp(p), _dp(v), _dp(o), _dp(n))
    return float(vnew), n


def generate_goals(formation: int, size: float, n: int, center, layer_dist: float) -> np.ndarray:
    """QuadrotorScenario.generate_goals of the oracle for formation index `formation` (QUADS_FORMATION_LIST order)."""
    out = np.zeros((max(n, 3) + 1, 3))
    c = np.ascontiguousarray(center, dtype=np.float64)
    rows = lib().qo_generate_goals(formation, float(size), int(n), _dp(c), float(layer_dist), cube_floor_dim(int(n)), _dp(out))
    return out[:rows].copy()


class OracleBatch:
    """N independent oracle envs stepped from `threads` host threads -- the CPU baseline of bench.py."""

    def __init__(self, cfg: QuadSimConfig, threads: int = 1):
        self.cfg, self.N, self.K = cfg, cfg.num_envs, cfg.num_agents
        self.envs = [OracleEnv(cfg, i) for i in range(self.N)]
        self.D = self.envs[0].D
        self.handles = (C.c_void_p * self.N)(*[e.h for e in self.envs])
        self.threads = max(1, min(threads, self.N))
        self.pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        bounds = np.linspace(0, self.N, self.threads + 1).astype(int)
        self.slices = [(int(bounds[i]), int(bounds[i + 1])) for i in range(self.threads)]
        self.obs = np.zeros((self.N * self.K, self.D))
        self.rew = np.zeros(self.N * self.K)
        self.done = np.zeros(self.N * self.K, dtype=np.uint8)

    def _sub(self, lo, hi):
        return C.cast(C.byref(self.handles, lo * C.sizeof(C.c_void_p)), C.POINTER(C.c_void_p))

    def reset(self):
        lib().qo_batch_reset(self.handles, self.N, _dp(self.obs))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.N * self.K, self.envs[0].A)
        K, D = self.K, self.D
        u8 = C.POINTER(C.c_uint8)

        def run(sl):
            lo, hi = sl
            lib().qo_batch_step(self._sub(lo, hi), hi - lo, _dp(a[lo * K:]), _dp(self.obs[lo * K:]),
                                _dp(self.rew[lo * K:]), self.done[lo * K:].ctypes.data_as(u8))

        if self.pool is None:
            run(self.slices[0])
        else:
            list(self.pool.map(run, self.slices))
        return self.obs, self.rew, self.done

    def set_ticks(self, ticks):
        """Episode clock of every env (bench.py staggers them so that resets are spread over the timed window)."""
        for e, t in zip(self.envs, ticks):
            lib().qo_set_tick(e.h, int(t))

    def run(self, action_pool, steps: int) -> int:
        """`steps` lock-step batch steps inside one native call on `self.threads` OpenMP threads (action set s % pool at
        step s).  Returns the number of episodes that finished.  This is what bench.py times as the CPU arm."""
        a = np.ascontiguousarray(action_pool, dtype=np.float64).reshape(-1, self.N * self.K * self.envs[0].A)
        return int(lib().qo_batch_run(self.handles, self.N, _dp(a), a.shape[0], int(steps), _dp(self.obs), _dp(self.rew),
                                      self.done.ctypes.data_as(C.POINTER(C.c_uint8)), self.threads))
