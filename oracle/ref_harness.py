"""Reference harness -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Imports the *unmodified* reference simulator from /root/reference behind a few
``sys.modules`` stubs so that it can be driven in this
container, and provides a "noise tape": every unit random draw the reference
consumes on the hot path (standard normals, U[0,1) uniforms, ``choice`` results)
is recorded in consumption order, so the C oracle (oracle/quadsim_oracle.c) can
replay the very same draws and be compared value-for-value.

The tape only works when numba's JIT is disabled (``NUMBA_DISABLE_JIT=1`` must be
set before numba is first imported): the reference's ``@njit`` kernels then run as
plain Python and their ``np.random`` calls hit the patched numpy functions.

This module is used by ``tests/golden/make_golden.py`` (which writes the committed
golden fixtures) and by nothing that runs on the GPU box: /root/reference does not
exist there.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("QS_REFERENCE_ROOT", "/root/reference")


# --------------------------------------------------------------------------------------
# import stubs
# --------------------------------------------------------------------------------------
def install_stubs():
    """Pre-populate sys.modules with the handful of third-party names the reference imports
    at module scope but which are absent here (gymnasium, bezier, sample_factory, pyglet scene)."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_qs_stub", False):
        return

    gym = types.ModuleType("gymnasium")
    gym._qs_stub = True

    class Env:
        metadata = {}

        def __init__(self, *a, **k):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

    class Box:
        def __init__(self, low, high, dtype=np.float32, shape=None):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = dtype

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")

    def np_random(seed=None):
        return np.random.default_rng(seed), seed

    seeding.np_random = np_random
    utils.seeding = seeding
    gym.Env, gym.Wrapper, gym.spaces, gym.utils = Env, Wrapper, spaces, utils
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.utils"] = utils
    sys.modules["gymnasium.utils.seeding"] = seeding

    # `bezier` is not installed here.  ep_rand_bezier.py only needs Curve(nodes, degree=2).evaluate_multi(pts); the stub
    # evaluates the curve the package documents, B(s) = sum_k C(n,k) (1-s)^(n-k) s^k P_k, so the scenario code itself runs
    # unmodified.
    bz = types.ModuleType("bezier")

    class Curve:
        def __init__(self, nodes, degree):
            self.nodes = np.asarray(nodes, dtype=np.float64)
            self.degree = int(degree)
            assert self.nodes.shape[1] == self.degree + 1

        def evaluate_multi(self, s_vals):
            from math import comb
            s = np.asarray(s_vals, dtype=np.float64)
            n = self.degree
            out = np.zeros((self.nodes.shape[0], len(s)))
            for k in range(n + 1):
                out += comb(n, k) * ((1.0 - s) ** (n - k)) * (s ** k) * self.nodes[:, k:k + 1]
            return out

    bz.Curve = Curve
    sys.modules["bezier"] = bz

    sf = types.ModuleType("sample_factory")
    sfu = types.ModuleType("sample_factory.utils")
    sfuu = types.ModuleType("sample_factory.utils.utils")
    sfuu.experiment_dir = lambda cfg=None: "/tmp"
    sf.utils, sfu.utils = sfu, sfuu
    sys.modules["sample_factory"] = sf
    sys.modules["sample_factory.utils"] = sfu
    sys.modules["sample_factory.utils.utils"] = sfuu

    viz = types.ModuleType("gym_art.quadrotor_multi.quadrotor_multi_visualization")
    viz.Quadrotor3DSceneMulti = object
    sys.modules["gym_art.quadrotor_multi.quadrotor_multi_visualization"] = viz

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


# --------------------------------------------------------------------------------------
# noise tape
# --------------------------------------------------------------------------------------
TAPE_NORMAL, TAPE_UNIFORM, TAPE_CHOICE = 0, 1, 2


class Tape:
    """Ordered record of unit draws.  kinds[i] in {NORMAL, UNIFORM, CHOICE}, vals[i] the value."""

    def __init__(self, seed=0):
        self.rs = np.random.RandomState(seed)
        self.kinds = []
        self.vals = []
        self.zero_normals = False      # if True every standard normal drawn is 0 (noise-free runs)

    def mark(self):
        return len(self.vals)

    def since(self, mark):
        return (np.asarray(self.kinds[mark:], dtype=np.int32),
                np.asarray(self.vals[mark:], dtype=np.float64))

    # unit draws -------------------------------------------------------------------
    def randn(self, n=None):
        m = 1 if n is None else int(np.prod(n))
        v = np.zeros(m) if self.zero_normals else self.rs.standard_normal(m)
        self.kinds.extend([TAPE_NORMAL] * m)
        self.vals.extend(v.tolist())
        return float(v[0]) if n is None else v.reshape(n)

    def rand(self, n=None):
        m = 1 if n is None else int(np.prod(n))
        v = self.rs.random_sample(m)
        self.kinds.extend([TAPE_UNIFORM] * m)
        self.vals.extend(v.tolist())
        return float(v[0]) if n is None else v.reshape(n)

    def choice_ids(self, n_pop, size):
        ids = self.rs.choice(n_pop, size=size, replace=False)
        self.kinds.extend([TAPE_CHOICE] * len(ids))
        self.vals.extend([float(i) for i in ids])
        return ids

    # numpy-compatible front ends ----------------------------------------------------
    def normal(self, loc=0.0, scale=1.0, size=None):
        return loc + scale * self.randn(size)

    def uniform(self, low=0.0, high=1.0, size=None):
        if size is None and (np.ndim(low) > 0 or np.ndim(high) > 0):
            size = np.broadcast(low, high).shape
        return low + (np.asarray(high) - np.asarray(low)) * self.rand(size)

    def randn_fn(self, *shape):
        return self.randn(shape if shape else None)

    def rand_fn(self, *shape):
        return self.rand(shape if shape else None)

    def randint(self, low, high=None, size=None):
        # legacy np.random.randint(low, high): one taped uniform, low + floor(u * (high - low)) on the truncated bounds
        if high is None:
            low, high = 0, low
        low, high = int(low), int(high)
        if size is not None:                                     # run_away.py:20: size=2 -> one taped uniform per element, in order
            return np.array([low + int(np.floor(self.rand() * (high - low))) for _ in range(int(np.prod(size)))]).reshape(size)
        return low + int(np.floor(self.rand() * (high - low)))

    def choice(self, a, size=None, replace=True, p=None):
        assert not replace and p is None
        pop = np.arange(a) if np.isscalar(a) else np.asarray(list(a))
        if size is None:
            return pop[self.choice_ids(len(pop), 1)[0]]
        return pop[self.choice_ids(len(pop), int(size))]


class TapeGenerator:
    """Stands in for the ``np.random.default_rng`` Generator the fork threads through the env."""

    def __init__(self, tape):
        self.t = tape

    def uniform(self, low=0.0, high=1.0, size=None):
        return self.t.uniform(low, high, size)

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        assert size is None
        return int(low + np.floor(self.t.rand() * (high - low)))

    def shuffle(self, x):
        # Fisher-Yates with taped uniforms (static_same_goal shuffles identical goals, so order is moot)
        n = len(x)
        for i in range(n - 1, 0, -1):
            j = int(np.floor(self.t.rand() * (i + 1)))
            x[[i, j]] = x[[j, i]]

    def random(self, size=None):
        return self.t.rand(size)


def install_tape(tape):
    """Patch every RNG entry point the reference's hot path uses so draws come from ``tape``."""
    assert os.environ.get("NUMBA_DISABLE_JIT") == "1", "tape needs NUMBA_DISABLE_JIT=1"
    install_stubs()
    import numpy.random as nr
    nr.normal = tape.normal
    nr.uniform = tape.uniform
    nr.randn = tape.randn_fn
    nr.rand = tape.rand_fn
    nr.choice = tape.choice
    nr.randint = tape.randint
    import numba as nb
    nb.random = types.SimpleNamespace(uniform=tape.uniform)
    import gym_art.quadrotor_multi.sensor_noise as sn
    sn.normal, sn.uniform = tape.normal, tape.uniform       # imported by name (sensor_noise.py:3-4)


# --------------------------------------------------------------------------------------
# env factories
# --------------------------------------------------------------------------------------
def make_upstream_env(num_agents=8, quads_mode="static_same_goal", obs_repr="xyz_vxyz_R_omega",
                      neighbor_visible_num=6, neighbor_obs_type="pos_vel", use_obstacles=False,
                      obst_density=0.2, obst_size=0.6, obst_spawn_area=(8.0, 8.0), use_downwash=False,
                      room_dims=(10.0, 10.0, 10.0), ep_time=15.0, sense_noise="default",
                      thrust_noise_ratio=0.05, rew_coeff=None, seed=0, tape=None,
                      collision_hitbox_radius=2.0, collision_falloff_radius=4.0, use_numba=True):
    """``quadrotor_multi.QuadrotorEnvMulti`` built with the kwargs of
    swarm_rl/env_wrappers/quad_utils.py:36-66 (make_quadrotor_env_multi)."""
    install_stubs()
    from gym_art.quadrotor_multi import quadrotor_multi as qm
    from gym_art.quadrotor_multi.scenarios import mix

    # reference bug shim: Scenario_o_* take 4 ctor args, create_scenario passes 5
    if not getattr(mix, "_qs_shimmed", False):
        orig_eval_create = mix.create_scenario

        def create_scenario(quads_mode, envs, num_agents, room_dims, rng):
            cls = getattr(mix, "Scenario_" + quads_mode)
            try:
                return cls(quads_mode, envs, num_agents, room_dims, rng)
            except TypeError:
                return cls(quads_mode, envs, num_agents, room_dims)

        mix.create_scenario = create_scenario
        qm.create_scenario = create_scenario
        mix._qs_shimmed = True

    if rew_coeff is None:
        rew_coeff = dict(pos=1.0, effort=0.05, spin=0.1, vel=0.0, crash=1.0, orient=1.0, yaw=0.0,
                         quadcol_bin=5.0, quadcol_bin_smooth_max=10.0, quadcol_bin_obst=5.0)
    cfg = types.SimpleNamespace(seed=seed)
    env = qm.QuadrotorEnvMulti(
        cfg=cfg, num_agents=num_agents, ep_time=ep_time, rew_coeff=rew_coeff, obs_repr=obs_repr,
        neighbor_visible_num=neighbor_visible_num, neighbor_obs_type=neighbor_obs_type,
        collision_hitbox_radius=collision_hitbox_radius, collision_falloff_radius=collision_falloff_radius,
        use_obstacles=use_obstacles, obst_density=obst_density, obst_size=obst_size,
        obst_spawn_area=list(obst_spawn_area), use_downwash=use_downwash, use_numba=use_numba,
        quads_mode=quads_mode, room_dims=tuple(room_dims), use_replay_buffer=False,
        quads_view_mode=["topdown"], quads_render=False,
        dynamics_params="Crazyflie", raw_control=True, raw_control_zero_middle=True,
        dynamics_randomize_every=None,
        dynamics_change=dict(noise=dict(thrust_noise_ratio=thrust_noise_ratio),
                             damp=dict(vel=0, omega_quadratic=0)),
        dyn_sampler_1=None, sense_noise=sense_noise, init_random_state=False)
    if os.environ.get("NUMBA_DISABLE_JIT") == "1":
        # with the JIT off, OUNoiseNumba is a plain Python class and keeps theta/sigma as float64; the jitclass
        # spec types them float32 (numba_utils.py:66-74), which is what the real (JIT-on) reference computes with
        for e in env.envs:
            n = e.dynamics.thrust_noise
            n.theta, n.sigma, n.mu = np.float32(n.theta), np.float32(n.sigma), np.float32(n.mu)
    if tape is not None:
        gen = TapeGenerator(tape)
        env.rng = gen
        env.scenario.rng = gen
        if hasattr(env.scenario, "scenario") and env.scenario.scenario is not None:
            env.scenario.scenario.rng = gen
        for e in env.envs:
            e.rng = gen
    return env


def snapshot(env):
    """Per-drone dynamics state of a reference env as a dict of float64 arrays."""
    d = [e.dynamics for e in env.envs]
    return dict(
        pos=np.array([x.pos for x in d], dtype=np.float64),
        vel=np.array([x.vel for x in d], dtype=np.float64),
        rot=np.array([x.rot for x in d], dtype=np.float64),
        omega=np.array([x.omega for x in d], dtype=np.float64),
        rot_damp=np.array([x.thrust_rot_damp for x in d], dtype=np.float64),
        cmds_damp=np.array([x.thrust_cmds_damp for x in d], dtype=np.float64),
        ou=np.array([np.array(x.thrust_noise.state) for x in d], dtype=np.float64),
        on_floor=np.array([bool(x.on_floor) for x in d]),
        crashed_floor=np.array([bool(x.crashed_floor) for x in d]),
        crashed_wall=np.array([bool(x.crashed_wall) for x in d]),
        crashed_ceiling=np.array([bool(x.crashed_ceiling) for x in d]),
        goal=np.array([e.goal for e in env.envs], dtype=np.float64),
        tick=np.array([e.tick for e in env.envs], dtype=np.int64),
        since_last_svd=np.array([x.since_last_svd for x in d], dtype=np.float64),
    )


# --------------------------------------------------------------------------------------
# fork env (swarm_rl/sb_train.py path): quadrotor_multi_rewards.QuadrotorEnvMulti(QuadrotorEnvConfig)
# --------------------------------------------------------------------------------------
def make_fork_env(tape=None, **overrides):
    """``quadrotor_multi_rewards.QuadrotorEnvMulti`` built from ``swarm_rl.global_cfg.QuadrotorEnvConfig`` exactly as
    ``SB3QuadrotorEnv._make_env`` does (swarm_rl/env_wrappers/sb3_quad_env.py:36-41).  ``overrides`` are dataclass
    field values (num_agents, episode_duration, obs_repr, neighbor_obs_type, neighbor_visible_num, room_dims, ...)."""
    install_stubs()
    from swarm_rl.global_cfg import QuadrotorEnvConfig
    from gym_art.quadrotor_multi import quadrotor_multi_rewards as qmr
    cfg = QuadrotorEnvConfig()
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    if cfg.seed is None:
        cfg.seed = 0
    env = qmr.QuadrotorEnvMulti(cfg)
    # reference bug: Scenario_dynamic_repulsive.pos starts as an *int* array (dynamic_repulsive.py:34), so
    # the first episode's evader position is garbage (nan cast to int).  From the second reset on it is float64.  The
    # fixtures start from the float state every later episode has.
    env.scenario.pos = np.zeros(2, dtype=np.float64)
    if os.environ.get("NUMBA_DISABLE_JIT") == "1":
        for e in env.envs:
            n = e.dynamics.thrust_noise
            n.theta, n.sigma, n.mu = np.float32(n.theta), np.float32(n.sigma), np.float32(n.mu)
    if tape is not None:
        gen = TapeGenerator(tape)
        env.rng = gen
        env.scenario.rng = gen
        for e in env.envs:
            e.rng = gen
            e.dynamics.rng = gen
    return env


def fork_snapshot(env):
    """dynamics snapshot + the fork's extra per-drone state (pre-controller angle, 12 PIDs) and the evader position."""
    s = snapshot(env)
    pid_state = []
    for e in env.envs:
        c = e.pre_controller
        row = []
        for ctl in (c.position_controller, c.velocity_controller, c.attitude_controller, c.rate_controller):
            for p in (ctl.pid_x, ctl.pid_y, ctl.pid_z):
                row += [p.last_error, p.integral]
        pid_state.append(row)
    s["pid"] = np.array(pid_state, dtype=np.float64)                       # [K, 24]
    s["angle"] = np.array([e.pre_controller.angle for e in env.envs], dtype=np.float64)
    s["ang_vel"] = np.array([e.pre_controller.angular_velocity for e in env.envs], dtype=np.float64)
    s["evader"] = np.array(env.scenario.pos, dtype=np.float64)
    s["heading"] = np.array(env.heading, dtype=np.float64)                 # refreshed in step only (quadrotor_multi_rewards.py:647)
    return s
