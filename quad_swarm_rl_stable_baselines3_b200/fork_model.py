"""Host-side constants of the fork's cascaded PID pre-controller and pursuit task (QS_MODE_FORK).

The reference builds these once per drone object in Python:
`gym_art/quadrotor_multi/Controller/MultirotorModel.py:9-66` (ModelParams: mass, J, allocation matrix),
`Controller/{Position,Velocity,Attitude,Rate}Controller.py` (gains, saturations, anti-windup limits),
`Controller/Mixer.py:31-65` (normalised pseudo-inverse allocation), `Controller/Controller.py:27-31`
and `scenarios/dynamic_repulsive.py:26-36`.  Here they become the POD block `qs_fork_config` (include/quadsim.h).
Nothing in this file runs per step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class ControllerModel:
    """ModelParams defaults (MultirotorModel.py:11-26)."""
    n_motors: int = 4
    g: float = 9.81
    mass: float = 0.028
    kf: float = 0.00000000125
    km: float = 0.0025
    prop_radius: float = 0.00015
    arm_length: float = 0.04596
    body_height: float = 0.003
    max_rpm: float = 13000.0
    min_rpm: float = 1170.0

    def inertia_diag(self) -> List[float]:
        a, bh, m = self.arm_length, self.body_height, self.mass
        lat = m * (3.0 * a * a + bh * bh) / 12.0                   # MultirotorModel.py:37-39
        return [lat, lat, m * a * a / 2.0]

    def allocation(self) -> np.ndarray:
        s = 0.707                                                   # MultirotorModel.py:42-53
        alloc = np.array([[-s, s, s, -s], [-s, s, -s, s], [-1.0, -1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 1.0]])
        alloc[0] *= self.arm_length * self.kf
        alloc[1] *= self.arm_length * self.kf
        alloc[2] *= self.km * (3.0 * self.prop_radius) * self.kf
        alloc[3] *= self.kf
        return alloc

    def mixer(self) -> np.ndarray:
        """Mixer.calculate_allocation (Mixer.py:31-65): right pseudo-inverse, roll/pitch columns normalised per motor,
        yaw column replaced by its sign, throttle column 1 (PX4-style)."""
        A = self.allocation()
        inv = A.T @ np.linalg.inv(A @ A.T)
        for i in range(self.n_motors):
            n = float(np.hypot(inv[i, 0], inv[i, 1]))
            if n > 0:
                inv[i, 0:2] /= n
            v = inv[i, 2]
            inv[i, 2] = 1.0 if v > 1e-2 else (-1.0 if v < -1e-2 else 0.0)
        inv[:, 3] = 1.0
        return inv


@dataclass
class ForkParams:
    substeps: int = 8                       # quadrotor_multi_rewards.py:633
    capture_radius: float = 3.0             # global_cfg.py:37 initial_capture_radius (0.2 if None, quadrotor_multi_rewards.py:208)
    rew_existence: float = -0.1             # quadrotor_multi_rewards.py:739-746
    rew_captor: float = 100.0
    rew_helper: float = 100.0
    max_angular_rate: float = math.pi * 80.0 / 180.0    # Controller.py:31
    chaser_speed: float = 0.2               # Controller.py:89
    evader_v_max: float = 0.5               # dynamic_repulsive.py:31
    evader_dt: float = 1.0 / 200.0          # :32
    evader_arena: float = 5.0               # :35
    spawn_ring: float = 0.5                 # :75
    evader_r_min: float = 2.0               # :79
    evader_r_span: float = 3.0
    rate_out_scale: float = 800.0           # RateController.py:84-86
    cam_focal_length: float = 0.035         # swarm_rl/global_cfg.py:13-17
    cam_num: int = 3
    cam_target_size: float = 0.2            # neighbour_size_cam
    cam_pixel_noise: float = 3.0
    cam_fov_deg: float = 70.0               # quadrotor_multi_rewards.py:286-287
    cam_resolution: float = 640.0
    model: ControllerModel = field(default_factory=ControllerModel)

    def pid_table(self) -> List[List[float]]:
        """[12][kp, kd, ki, saturation, antiwindup] in cascade order."""
        J = self.model.inertia_diag()
        pos = [4.1625, 0.5473, 0.0023]      # PositionController.py:13-15
        vel = [2.4531, 0.0003, 0.0382]      # VelocityController.py:19-21
        att = [11.2081, 0.0490, 0.0073]     # AttitudeController.py:11-13
        rate = [3.1222, 0.0477, 0.0001]     # RateController.py:13-15
        t = []
        t += [pos + [6.0, 1.0], pos + [6.0, 1.0], pos + [6.0, 2.0]]           # PositionController.py:50-56
        t += [vel + [40.0, 1.0]] * 3                                          # VelocityController.py:57-62
        t += [att + [10.0, 0.1], att + [10.0, 0.1], att + [1.0, 0.1]]         # AttitudeController.py:49-54
        t += [[g * J[a] for g in rate] + [-1.0, 1.0] for a in range(3)]       # RateController.py:49-65
        return [list(map(float, r)) for r in t]
