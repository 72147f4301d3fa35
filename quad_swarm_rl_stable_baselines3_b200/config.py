"""Configuration: the reference's constructor kwargs flattened to the POD `qs_config` of include/quadsim.h.

`QuadSimConfig` takes the same names as `QuadrotorEnvMulti.__init__`
(gym_art/quadrotor_multi/quadrotor_multi.py:27-45) and the env factory
`make_quadrotor_env_multi` (swarm_rl/env_wrappers/quad_utils.py:20-66) where they exist.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

from . import fork_model, quad_model

QS_API_VERSION = 2
QS_MAX_AGENTS = 32
QS_MAX_OBSTACLES = 64

SCENARIOS = {"static_same_goal": 0, "o_mix": 1, "o_random": 2, "o_static_same_goal": 3, "dynamic_repulsive": 4,
             "static_diff_goal": 5, "dynamic_same_goal": 6, "dynamic_diff_goal": 7, "swap_goals": 8, "dynamic_formations": 9,
             "mix": 10, "ep_lissajous3D": 11, "ep_rand_bezier": 12, "swarm_vs_swarm": 13, "run_away": 14}
FORMATION_SCENARIOS = ("static_same_goal", "static_diff_goal", "dynamic_same_goal", "dynamic_diff_goal", "swap_goals",
                       "dynamic_formations", "mix", "ep_lissajous3D", "ep_rand_bezier", "swarm_vs_swarm", "run_away")
QS_SC_COUNT = 24       # floats per env of formation-scenario state (include/quadsim.h QS_SC_*)
SCENARIO_NAMES = {v: k for k, v in SCENARIOS.items()}
# qs_episode_records layout (include/quadsim.h QS_ER_*)
QS_ER_COUNT = 20
ER = dict(seq=0, scenario=1, num_collisions=2, num_collisions_after_settle=3, num_collisions_final_5_s=4,
          num_collisions_with_room=5, num_collisions_with_floor=6, num_collisions_with_wall=7, num_collisions_with_ceiling=8,
          num_collisions_obst_quad=9, num_collisions_obst_quad_after_settle=10, agents_success=11, agents_deadlock=12,
          agents_collided=13, agents_neighbor_col=14, agents_obst_col=15, ep_len=16, success=17, nonfinite=18)


def episode_extra_stats(env_rec, agent_rec, num_agents: int, use_obstacles: bool) -> dict:
    """infos[i]['episode_extra_stats'] of the reference (quadrotor_multi.py:739-831, quadrotor_multi_rewards.py:886-978) for one
    agent, from the env's record row (`env_rec`, QS_ER_COUNT ints) and that agent's row of the per-drone part (`agent_rec`)."""
    name = SCENARIO_NAMES.get(int(env_rec[ER["scenario"]]), "unknown")
    if name == "o_mix":
        name = "mix"
    K = float(num_agents)
    d1, d3, d5 = (float(agent_rec[0]), float(agent_rec[1]), float(agent_rec[2]))
    g = lambda k: int(env_rec[ER[k]])  # noqa: E731
    s = {
        "num_collisions": g("num_collisions"),
        "num_collisions_with_room": g("num_collisions_with_room"),
        "num_collisions_with_floor": g("num_collisions_with_floor"),
        "num_collisions_with_wall": g("num_collisions_with_wall"),
        "num_collisions_with_ceiling": g("num_collisions_with_ceiling"),
        "num_collisions_after_settle": g("num_collisions_after_settle"),
        f"{name}/num_collisions": g("num_collisions_after_settle"),
        "num_collisions_final_5_s": g("num_collisions_final_5_s"),
        f"{name}/num_collisions_final_5_s": g("num_collisions_final_5_s"),
        "distance_to_goal_1s": d1, "distance_to_goal_3s": d3, "distance_to_goal_5s": d5,
        f"{name}/distance_to_goal_1s": d1, f"{name}/distance_to_goal_3s": d3, f"{name}/distance_to_goal_5s": d5,
    }
    if use_obstacles:
        s["num_collisions_obst_quad"] = g("num_collisions_obst_quad")
        s["num_collisions_obst_quad_after_settle"] = g("num_collisions_obst_quad_after_settle")
        s[f"{name}/num_collisions_obst"] = g("num_collisions_obst_quad")
        # distance_to_goal_3_5 / distance_to_goal_5 are zeroed at reset and never updated in the reference (:493-494)
        s["num_collisions_obst_quad_3_5"] = s[f"{name}/num_collisions_obst_quad_3_5"] = 0
        s["num_collisions_obst_quad_5"] = s[f"{name}/num_collisions_obst_quad_5"] = 0
    for key, col in (("agent_success_rate", "agents_success"), ("agent_deadlock_rate", "agents_deadlock"),
                     ("agent_col_rate", "agents_collided"), ("agent_neighbor_col_rate", "agents_neighbor_col"),
                     ("agent_obst_col_rate", "agents_obst_col")):
        s["metric/" + key] = s[f"{name}/{key}"] = g(col) / K
    return s


ENV_MODES = {"upstream": 0, "fork": 1}
OBS_REPR = {"xyz_vxyz_R_omega": 0, "xyz_vxyz_R_omega_floor": 1, "xyz_vxyz_R_omega_wall": 2,
            "cdist_cdistdot_dist_distdot_angle_angledot": 3, "cdist_cdistdot_dist_distdot_sangle_angledot": 4,
            "aw_awdot_dist_distdot_angle_angledot": 5, "cdist_cdistdot_ndist_distdot_nsangle_angledot": 6}
OBS_REPR_DIM = {"xyz_vxyz_R_omega": 18, "xyz_vxyz_R_omega_floor": 19, "xyz_vxyz_R_omega_wall": 24,       # quad_utils.py:30-38
                "cdist_cdistdot_dist_distdot_angle_angledot": 6, "cdist_cdistdot_dist_distdot_sangle_angledot": 7,
                "aw_awdot_dist_distdot_angle_angledot": 6, "cdist_cdistdot_ndist_distdot_nsangle_angledot": 7}
FORK_OBS_REPR = {k for k, v in OBS_REPR.items() if v >= 3}
NEIGHBOR_OBS = {"none": 0, "pos_vel": 1, "dist_angle": 2, "dist_sangle": 3, "dist_angle_heading": 4,
                "dist_sangle_sheading": 5, "ndist_nsangle": 6}
NEIGHBOR_OBS_DIM = {"none": 0, "pos_vel": 6, "dist_angle": 2, "dist_sangle": 3, "dist_angle_heading": 3,      # quad_utils.py:40-58
                    "dist_sangle_sheading": 5, "ndist_nsangle": 3}
FORK_NEIGHBOR_OBS = {"none", "dist_angle", "dist_sangle", "dist_angle_heading", "dist_sangle_sheading", "ndist_nsangle"}
PARAM_KEYS = {"pos": 0, "effort": 1, "crash": 2, "orient": 3, "spin": 4,
              "quadcol_bin": 5, "quadcol_bin_smooth_max": 6, "quadcol_bin_obst": 7, "capture_radius": 8}


def cube_floor_dim(n: int) -> int:
    """floor_dim_size of the cube formation exactly as the reference evaluates it (scenarios/base.py:101-102)."""
    import numpy as np
    return int(np.power(n, 1.0 / 3)) if n > 0 else 1


class QsForkConfigC(C.Structure):
    """Binary mirror of `struct qs_fork_config`."""
    _fields_ = [("substeps", C.c_int32), ("reserved", C.c_int32)] + [(n, C.c_double) for n in (
        "capture_radius", "rew_existence", "rew_captor", "rew_helper", "max_angular_rate", "chaser_speed",
        "evader_v_max", "evader_dt", "evader_arena", "spawn_ring", "evader_r_min", "evader_r_span")] + [
        ("pid", (C.c_double * 5) * 12), ("rate_out_scale", C.c_double), ("mixer", (C.c_double * 4) * 4),
        ("ctrl_mass", C.c_double), ("ctrl_g", C.c_double), ("ctrl_kf", C.c_double), ("ctrl_min_rpm", C.c_double),
        ("ctrl_max_rpm", C.c_double), ("cam_focal_length", C.c_double), ("cam_target_size", C.c_double),
        ("cam_pixel_noise", C.c_double), ("cam_fov_deg", C.c_double), ("cam_resolution", C.c_double),
        ("cam_num", C.c_int32), ("reserved2", C.c_int32)]


class QsConfigC(C.Structure):
    """Binary mirror of `struct qs_config` (include/quadsim.h).  tests/test_capi_cpu.py checks sizeof agrees."""
    _fields_ = [
        ("api_version", C.c_int32), ("num_envs", C.c_int32), ("num_agents", C.c_int32), ("scenario", C.c_int32),
        ("obs_repr", C.c_int32), ("neighbor_obs_type", C.c_int32), ("neighbor_visible_num", C.c_int32),
        ("use_obstacles", C.c_int32), ("use_downwash", C.c_int32), ("apply_collision_force", C.c_int32),
        ("sense_noise", C.c_int32), ("ep_len", C.c_int32), ("sim_steps", C.c_int32), ("svd_period", C.c_int32),
        ("obst_area_len", C.c_int32), ("obst_area_wid", C.c_int32), ("num_obstacles", C.c_int32),
        ("env_mode", C.c_int32), ("cube_dim", C.c_int32 * 3), ("reserved0", C.c_int32),
        ("seed", C.c_uint64), ("env_id_offset", C.c_int64),
        ("dt", C.c_double), ("room_dims", C.c_double * 3), ("gravity", C.c_double),
        ("mass", C.c_double), ("inertia", C.c_double * 3), ("thrust_max", C.c_double * 4),
        ("torque_max", C.c_double * 4), ("prop_cross", (C.c_double * 3) * 4), ("prop_ccw", C.c_double * 4),
        ("arm", C.c_double), ("motor_tau_up", C.c_double), ("motor_tau_down", C.c_double),
        ("motor_linearity", C.c_double), ("vel_damp", C.c_double), ("damp_omega_quadratic", C.c_double),
        ("omega_max", C.c_double), ("floor_mu", C.c_double), ("ou_theta", C.c_double), ("ou_sigma", C.c_double),
        ("sense_pos_std", C.c_double), ("sense_vel_std", C.c_double), ("sense_gyro_std", C.c_double),
        ("rew_pos", C.c_double), ("rew_effort", C.c_double), ("rew_crash", C.c_double), ("rew_orient", C.c_double),
        ("rew_spin", C.c_double), ("rew_quadcol_bin", C.c_double), ("rew_quadcol_bin_smooth_max", C.c_double),
        ("rew_quadcol_bin_obst", C.c_double),
        ("collision_hitbox_radius", C.c_double), ("collision_falloff_radius", C.c_double),
        ("spawn_box", C.c_double), ("spawn_min_z", C.c_double),
        ("obst_size", C.c_double), ("sdf_resolution", C.c_double), ("approach_goal_metric", C.c_double),
        ("fork", QsForkConfigC),
    ]


class QsStatsC(C.Structure):
    """Binary mirror of `struct qs_stats`."""
    _fields_ = [(n, C.c_int64) for n in (
        "episodes", "num_collisions", "num_collisions_after_settle", "num_collisions_final_5s",
        "num_collisions_with_room", "num_collisions_with_floor", "num_collisions_with_wall",
        "num_collisions_with_ceiling", "num_collisions_obst_quad", "num_collisions_obst_quad_after_settle",
        "agents_success", "agents_deadlock", "agents_collided", "nonfinite_resets", "episodes_success")] + [
        (n, C.c_double) for n in ("distance_to_goal_1s", "distance_to_goal_3s", "distance_to_goal_5s", "reward_sum")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class QsStateViewC(C.Structure):
    """Binary mirror of `struct qs_state_view`: raw device addresses."""
    FIELDS = ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "ou", "goal", "flags", "col_mask",
              "tick", "svd_ctr", "step_ctr", "obst_xy", "scenario", "pid", "heading", "evader")
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


# default reward coefficients: quadrotor_multi.py:101-104 overlaid with the upstream training recipe
# (swarm_rl/runs/quad_multi_mix_baseline.py:13-16: collision reward 5, smooth max penalty 10)
DEFAULT_REW_COEFF = dict(pos=1.0, effort=0.05, crash=1.0, orient=1.0, spin=0.1,
                         quadcol_bin=5.0, quadcol_bin_smooth_max=10.0, quadcol_bin_obst=5.0)


@dataclass
class QuadSimConfig:
    num_envs: int = 1
    num_agents: int = 8                       # quads_num_agents
    quads_mode: str = "static_same_goal"      # 'static_same_goal' | 'mix' (with obstacles) | 'o_random' | 'o_static_same_goal'
    obs_repr: str = "xyz_vxyz_R_omega"
    neighbor_obs_type: str = "pos_vel"
    neighbor_visible_num: int = 6             # -1 = all others
    collision_hitbox_radius: float = 2.0
    collision_falloff_radius: float = 4.0
    use_obstacles: bool = False
    obst_density: float = 0.2
    obst_size: float = 0.6
    obst_spawn_area: Sequence[float] = (8.0, 8.0)
    use_downwash: bool = False
    apply_collision_force: bool = True
    room_dims: Sequence[float] = (10.0, 10.0, 10.0)
    ep_time: float = 15.0                     # quads_episode_duration
    sim_freq: float = 200.0
    sim_steps: int = 2
    sense_noise: Optional[str] = "default"    # None | 'default'
    rew_coeff: Dict[str, float] = field(default_factory=dict)
    seed: int = 0
    env_id_offset: int = 0
    motor: quad_model.MotorParams = field(default_factory=quad_model.MotorParams)
    geometry: quad_model.QuadGeometry = field(default_factory=quad_model.QuadGeometry)
    env_mode: str = "upstream"                # 'upstream' (quadrotor_multi.py) | 'fork' (quadrotor_multi_rewards.py)
    fork: fork_model.ForkParams = field(default_factory=fork_model.ForkParams)

    @classmethod
    def fork_default(cls, **kw) -> "QuadSimConfig":
        """The env `swarm_rl/sb_train.py` trains on: defaults of `swarm_rl/global_cfg.py:QuadrotorEnvConfig`
        (4 agents, dynamic_repulsive, 2-D observations, room 15x15x3, 30 s episodes, collision forces off)."""
        base = dict(env_mode="fork", num_agents=4, quads_mode="dynamic_repulsive",
                    obs_repr="cdist_cdistdot_dist_distdot_angle_angledot", neighbor_obs_type="dist_angle",
                    neighbor_visible_num=-1, room_dims=(15.0, 15.0, 3.0), ep_time=30.0,
                    apply_collision_force=False)                  # quadrotor_multi_rewards.py:203
        capture_radius = kw.pop("capture_radius", None)
        camera = kw.pop("camera", None)
        base.update(kw)
        cfg = cls(**base)
        if capture_radius is not None:
            cfg.fork.capture_radius = float(capture_radius)
        for k, v in (camera or {}).items():          # cam_focal_length / cam_target_size / cam_pixel_noise / cam_num
            if not hasattr(cfg.fork, k):
                raise KeyError(k)
            setattr(cfg.fork, k, v)
        return cfg

    @classmethod
    def from_reference_cfg(cls, rcfg, num_envs: int, **kw) -> "QuadSimConfig":
        """Map a `swarm_rl.global_cfg.QuadrotorEnvConfig` (duck-typed) to the fork-mode configuration."""
        cap = getattr(rcfg, "initial_capture_radius", None)
        return cls.fork_default(
            num_envs=num_envs, num_agents=rcfg.num_agents, quads_mode=rcfg.quads_mode, obs_repr=rcfg.obs_repr,
            neighbor_obs_type=rcfg.neighbor_obs_type, neighbor_visible_num=rcfg.neighbor_visible_num,
            room_dims=tuple(float(v) for v in rcfg.room_dims), ep_time=float(rcfg.episode_duration),
            collision_hitbox_radius=rcfg.collision_hitbox_radius, collision_falloff_radius=rcfg.collision_falloff_radius,
            use_downwash=bool(rcfg.use_downwash), sim_freq=float(rcfg.sim_freq), sim_steps=int(rcfg.sim_steps),
            sense_noise=rcfg.sense_noise, seed=int(rcfg.seed or 0),
            capture_radius=0.2 if cap is None else float(cap),
            camera=dict(cam_focal_length=float(rcfg.focal_length_cam), cam_target_size=float(rcfg.neighbour_size_cam),
                        cam_pixel_noise=float(rcfg.pixel_noise_cam), cam_num=int(rcfg.n_cameras)), **kw)

    # ---- derived ---------------------------------------------------------------------------
    @property
    def visible(self) -> int:
        if self.neighbor_obs_type == "none":
            return 0
        v = self.num_agents - 1 if self.neighbor_visible_num == -1 else self.neighbor_visible_num
        if not 0 <= v <= self.num_agents - 1:
            raise ValueError("Incorrect number of neigbors")      # quadrotor_multi.py:375
        return v

    @property
    def obs_dim(self) -> int:
        return (OBS_REPR_DIM[self.obs_repr] + NEIGHBOR_OBS_DIM[self.neighbor_obs_type] * self.visible
                + (9 if self.use_obstacles else 0))

    @property
    def act_dim(self) -> int:
        return 2 if self.env_mode == "fork" else 4      # quadrotor_control.py:74-86 (CustomPidControl) vs :37-49

    @property
    def dt(self) -> float:
        return 1.0 / self.sim_freq

    @property
    def ep_len(self) -> int:
        return int(self.ep_time / (self.dt * self.sim_steps))     # quadrotor_single.py:158

    @property
    def num_obstacles(self) -> int:
        if not self.use_obstacles:
            return 0
        return int(self.obst_density * self.obst_spawn_area[0] * self.obst_spawn_area[1])  # quadrotor_multi.py:138

    def scenario_id(self) -> int:
        mode = self.quads_mode
        if self.use_obstacles and mode == "mix":
            mode = "o_mix"
        if mode not in SCENARIOS:
            raise ValueError(f"quads_mode {self.quads_mode!r} is not available on the device (supported without obstacles: "
                             f"{', '.join(FORMATION_SCENARIOS)}; with obstacles: mix, o_random, o_static_same_goal; "
                             f"fork mode: dynamic_repulsive)")
        if self.use_obstacles != mode.startswith("o_"):
            raise ValueError(f"quads_mode {self.quads_mode!r} inconsistent with use_obstacles={self.use_obstacles}")
        if (mode == "dynamic_repulsive") != (self.env_mode == "fork"):
            raise ValueError("dynamic_repulsive is the fork-mode scenario (env_mode='fork') and the only one it supports")
        K = self.num_agents
        # a sphere of fewer than 3 drones still has 3 goal rows in the reference (scenarios/utils.py:77-80); scenarios that
        # permute stored rows need every row to belong to a drone
        if mode == "swap_goals" and K < 3:
            raise ValueError("swap_goals needs num_agents >= 3 on the device")
        if mode == "swarm_vs_swarm" and K < 2:
            raise ValueError("swarm_vs_swarm needs num_agents >= 2")             # scenarios/utils.py:10
        if mode == "run_away" and K < 2:
            raise ValueError("run_away needs num_agents >= 2")                   # run_away.py:20,23-24: randint(1, K), envs[1]
        if mode == "mix" and K == 2:
            raise ValueError("mix needs num_agents == 1 or >= 3 on the device (it contains swap_goals)")
        return SCENARIOS[mode]

    def to_c(self) -> QsConfigC:
        if not 1 <= self.num_agents <= QS_MAX_AGENTS:
            raise ValueError(f"num_agents must be in [1, {QS_MAX_AGENTS}]")
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        if self.env_mode not in ENV_MODES:
            raise ValueError(f"env_mode {self.env_mode!r} not supported")
        if self.obs_repr not in OBS_REPR:
            raise ValueError(f"obs_repr {self.obs_repr!r} not supported")
        if self.neighbor_obs_type not in NEIGHBOR_OBS:
            raise ValueError(f"neighbor_obs_type {self.neighbor_obs_type!r} not supported")
        is_fork = self.env_mode == "fork"
        if (self.obs_repr in FORK_OBS_REPR) != is_fork:
            raise ValueError(f"obs_repr {self.obs_repr!r} does not belong to env_mode {self.env_mode!r}")
        if is_fork and self.neighbor_obs_type not in FORK_NEIGHBOR_OBS:
            raise ValueError(f"neighbor_obs_type {self.neighbor_obs_type!r} is not a fork-mode type")
        if not is_fork and self.neighbor_obs_type in FORK_NEIGHBOR_OBS - {"none"}:
            raise ValueError(f"neighbor_obs_type {self.neighbor_obs_type!r} needs env_mode='fork'")
        if is_fork and self.use_obstacles:
            raise ValueError("the fork env has its obstacle path commented out (quadrotor_multi_rewards.py:106-119)")
        rew = dict(DEFAULT_REW_COEFF)
        unknown = set(self.rew_coeff) - set(rew) - {"action_change", "yaw", "rot", "attitude", "vel"}
        if unknown:
            raise AssertionError(f"unknown rew_coeff keys {sorted(unknown)}")      # quadrotor_multi.py:109,116
        rew.update({k: float(v) for k, v in self.rew_coeff.items() if k in rew})
        q = quad_model.crazyflie_constants(self.geometry, self.motor, self.dt)
        c = QsConfigC()
        c.api_version = QS_API_VERSION
        c.env_mode = ENV_MODES[self.env_mode]
        c.num_envs, c.num_agents = self.num_envs, self.num_agents
        c.scenario = self.scenario_id()
        K = self.num_agents
        for i, n in enumerate((K, K // 2, K - K // 2)):
            c.cube_dim[i] = cube_floor_dim(n)
        c.obs_repr = OBS_REPR[self.obs_repr]
        c.neighbor_obs_type = NEIGHBOR_OBS[self.neighbor_obs_type] if self.visible > 0 else 0
        if is_fork and self.fork.substeps < 1:
            raise ValueError("fork.substeps must be >= 1")
        c.neighbor_visible_num = self.visible
        c.use_obstacles = int(self.use_obstacles)
        c.use_downwash = int(self.use_downwash)
        c.apply_collision_force = int(self.apply_collision_force)
        c.sense_noise = 0 if self.sense_noise is None else 1
        if self.sense_noise not in (None, "default"):
            raise ValueError("sense_noise must be None or 'default'")
        c.ep_len, c.sim_steps = self.ep_len, self.sim_steps
        c.svd_period = quad_model.svd_period(self.dt)
        c.obst_area_len, c.obst_area_wid = int(self.obst_spawn_area[0]), int(self.obst_spawn_area[1])
        c.num_obstacles = self.num_obstacles
        if c.num_obstacles > QS_MAX_OBSTACLES or c.obst_area_len * c.obst_area_wid > 64:
            raise ValueError("obstacle grid too large (<= 64 cells, <= 64 obstacles)")
        if self.use_obstacles and c.obst_area_len * c.obst_area_wid - c.num_obstacles < self.num_agents:
            raise ValueError("not enough free cells to spawn every drone in its own cell")
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.env_id_offset = self.env_id_offset
        c.dt = self.dt
        c.room_dims[:] = [float(v) for v in self.room_dims]
        c.gravity = quad_model.GRAV
        c.mass = q.mass
        c.inertia[:] = q.inertia
        c.thrust_max[:] = q.thrust_max
        c.torque_max[:] = q.torque_max
        for i in range(4):
            c.prop_cross[i][:] = q.prop_cross[i]
        c.prop_ccw[:] = q.prop_ccw
        c.arm = q.arm
        c.motor_tau_up, c.motor_tau_down = q.motor_tau_up, q.motor_tau_down
        c.motor_linearity = q.motor_linearity
        c.vel_damp, c.damp_omega_quadratic = q.vel_damp, q.damp_omega_quadratic
        c.omega_max, c.floor_mu = 40.0, 0.6                       # quadrotor_dynamics.py:49,77
        c.ou_theta, c.ou_sigma = 0.15, q.ou_sigma                 # numba_utils.py:80, quadrotor_dynamics.py:173
        c.sense_pos_std, c.sense_vel_std, c.sense_gyro_std = 0.005, 0.01, 0.000175   # sensor_noise.py:70-74
        c.rew_pos, c.rew_effort, c.rew_crash = rew["pos"], rew["effort"], rew["crash"]
        c.rew_orient, c.rew_spin = rew["orient"], rew["spin"]
        c.rew_quadcol_bin = rew["quadcol_bin"]
        c.rew_quadcol_bin_smooth_max = rew["quadcol_bin_smooth_max"]
        c.rew_quadcol_bin_obst = rew["quadcol_bin_obst"]
        c.collision_hitbox_radius = self.collision_hitbox_radius
        c.collision_falloff_radius = self.collision_falloff_radius
        c.spawn_box = 0.1 if self.use_obstacles else 2.0          # quadrotor_single.py:215-218
        c.spawn_min_z = 0.75                                      # quadrotor_single.py:416-417
        c.obst_size = self.obst_size
        c.sdf_resolution = 0.1                                    # obstacles/obstacles.py:13
        c.approach_goal_metric = 0.5                              # scenarios/base.py:35
        f, fp = c.fork, self.fork
        f.substeps = fp.substeps
        for n in ("capture_radius", "rew_existence", "rew_captor", "rew_helper", "max_angular_rate", "chaser_speed",
                  "evader_v_max", "evader_dt", "evader_arena", "spawn_ring", "evader_r_min", "evader_r_span",
                  "rate_out_scale"):
            setattr(f, n, float(getattr(fp, n)))
        for i, row in enumerate(fp.pid_table()):
            f.pid[i][:] = row
        mx = fp.model.mixer()
        for i in range(4):
            f.mixer[i][:] = [float(v) for v in mx[i]]
        f.ctrl_mass, f.ctrl_g, f.ctrl_kf = fp.model.mass, fp.model.g, fp.model.kf
        f.ctrl_min_rpm, f.ctrl_max_rpm = fp.model.min_rpm, fp.model.max_rpm
        f.cam_focal_length, f.cam_target_size, f.cam_pixel_noise = fp.cam_focal_length, fp.cam_target_size, fp.cam_pixel_noise
        f.cam_fov_deg, f.cam_resolution, f.cam_num = fp.cam_fov_deg, fp.cam_resolution, int(fp.cam_num)
        return c
