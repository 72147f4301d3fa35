"""Host-side quadrotor model constants.

The reference evaluates a pile of construction-time Python once per drone object
(`gym_art/quadrotor_multi/quad_models.py:1-43` parameter dicts,
`gym_art/quadrotor_multi/inertia.py:182-310` link-set inertia,
`gym_art/quadrotor_multi/quadrotor_dynamics.py:106-168` derived motor constants).
Here that becomes one small constant block computed on the host and shipped to the device
inside `qs_config` (include/quadsim.h).  Nothing in this file runs per step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

GRAV = 9.81


@dataclass
class Box:
    l: float
    w: float
    h: float
    m: float

    def inertia_diag(self):
        # solid cuboid about its own centre, axes x=l, y=w, z=h (inertia.py:90-96)
        return (self.m * (self.h ** 2 + self.w ** 2) / 12.0,
                self.m * (self.l ** 2 + self.h ** 2) / 12.0,
                self.m * (self.w ** 2 + self.l ** 2) / 12.0)


@dataclass
class Cylinder:
    h: float
    r: float
    m: float

    def inertia_diag(self):
        # vertical cylinder (inertia.py:149-155)
        lat = self.m * (3 * self.r ** 2 + self.h ** 2) / 12.0
        return (lat, lat, 0.5 * self.m * self.r ** 2)


@dataclass
class QuadGeometry:
    """Crazyflie 2.x geometry (quad_models.py:4-15)."""
    body: Box = field(default_factory=lambda: Box(0.03, 0.03, 0.004, 0.005))
    payload: Box = field(default_factory=lambda: Box(0.035, 0.02, 0.008, 0.01))
    arm: Box = field(default_factory=lambda: Box(0.022, 0.005, 0.005, 0.001))
    motor: Cylinder = field(default_factory=lambda: Cylinder(0.02, 0.0035, 0.0015))
    propeller: Cylinder = field(default_factory=lambda: Cylinder(0.002, 0.022, 0.00075))
    motor_xyz: Sequence[float] = (0.065 / 2, 0.065 / 2, 0.0)
    arm_angle_deg: float = 45.0
    arm_z: float = 0.0
    payload_xy: Sequence[float] = (0.0, 0.0)
    payload_z_sign: float = 1.0


@dataclass
class MotorParams:
    """quad_models.py:25-33 with the env factory's `dynamics_change` overlay
    (swarm_rl/env_wrappers/quad_utils.py:33: thrust_noise_ratio 0.05, zero damping)."""
    thrust_to_weight: float = 1.9
    assymetry: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    torque_to_thrust: float = 0.006
    linearity: float = 1.0
    damp_time_up: float = 0.15
    damp_time_down: float = 0.15
    thrust_noise_ratio: float = 0.05
    vel_damp: float = 0.0
    omega_quadratic_damp: float = 0.0


@dataclass
class QuadConstants:
    mass: float
    inertia: List[float]
    thrust_max: List[float]
    torque_max: List[float]
    prop_pos: List[List[float]]
    prop_cross: List[List[float]]
    prop_ccw: List[float]
    arm: float
    motor_tau_up: float
    motor_tau_down: float
    motor_linearity: float
    vel_damp: float
    damp_omega_quadratic: float
    ou_sigma: float


def _yaw_rotated_diag(diag, alpha):
    """diag(R I R^T) for a rotation by alpha about z of a diagonal tensor (inertia.py:15-20)."""
    c2, s2 = math.cos(alpha) ** 2, math.sin(alpha) ** 2
    return (diag[0] * c2 + diag[1] * s2, diag[0] * s2 + diag[1] * c2, diag[2])


def crazyflie_constants(geom: QuadGeometry | None = None, motor: MotorParams | None = None,
                        dt: float = 1.0 / 200.0) -> QuadConstants:
    geom = geom or QuadGeometry()
    motor = motor or MotorParams()
    a = math.radians(geom.arm_angle_deg) or 0.01
    mx, my, mz = geom.motor_xyz
    delta_y = my - geom.body.w / 2.0
    arm_xyz = (mx - delta_y / (2.0 * math.tan(a)), my - delta_y / 2.0, geom.arm_z)
    # bodies are listed clockwise from front-right (inertia.py:236-240)
    xs, ys = (1, -1, -1, 1), (-1, -1, 1, 1)
    arm_yaw = (-a, a, -a, a)
    prop_z = mz + geom.motor.h / 2.0 + geom.propeller.h

    # (link, centre xyz, yaw)
    links = [(geom.body, (0.0, 0.0, 0.0), 0.0),
             (geom.payload, (geom.payload_xy[0], geom.payload_xy[1],
                             math.copysign(1.0, geom.payload_z_sign) * (geom.body.h + geom.payload.h) / 2.0), 0.0)]
    links += [(geom.arm, (xs[i] * arm_xyz[0], ys[i] * arm_xyz[1], arm_xyz[2]), arm_yaw[i]) for i in range(4)]
    links += [(geom.motor, (xs[i] * mx, ys[i] * my, mz), 0.0) for i in range(4)]
    links += [(geom.propeller, (xs[i] * mx, ys[i] * my, prop_z), 0.0) for i in range(4)]

    mass = sum(l.m for l, _, _ in links)
    com = [sum(l.m * p[k] for l, p, _ in links) / mass for k in range(3)]
    inertia = [0.0, 0.0, 0.0]
    for link, p, yaw in links:
        d = _yaw_rotated_diag(link.inertia_diag(), yaw)
        x, y, z = (p[0] - com[0], p[1] - com[1], p[2] - com[2])
        inertia[0] += d[0] + link.m * (y * y + z * z)       # parallel axis (inertia.py:22-36)
        inertia[1] += d[1] + link.m * (x * x + z * z)
        inertia[2] += d[2] + link.m * (x * x + y * y)

    prop_pos = [[xs[i] * mx - com[0], ys[i] * my - com[1], mz - com[2]] for i in range(4)]
    prop_cross = [[p[1], -p[0], 0.0] for p in prop_pos]       # cross(p, z^)  (quadrotor_dynamics.py:143)
    asym = [v * 4.0 / sum(motor.assymetry) for v in motor.assymetry]
    thrust_max = [GRAV * mass * motor.thrust_to_weight * v / 4.0 for v in asym]
    torque_max = [motor.torque_to_thrust * t for t in thrust_max]
    return QuadConstants(
        mass=mass, inertia=inertia, thrust_max=thrust_max, torque_max=torque_max, prop_pos=prop_pos,
        prop_cross=prop_cross, prop_ccw=[-1.0, 1.0, -1.0, 1.0], arm=math.hypot(mx, my),
        motor_tau_up=4 * dt / (motor.damp_time_up + 1e-6), motor_tau_down=4 * dt / (motor.damp_time_down + 1e-6),
        motor_linearity=motor.linearity, vel_damp=motor.vel_damp, damp_omega_quadratic=motor.omega_quadratic_damp,
        ou_sigma=0.2 * motor.thrust_noise_ratio)


def svd_period(dt: float, limit: float = 0.5) -> int:
    """Sub-steps between re-orthonormalisations: the reference accumulates `since_last_svd += dt` in float64 and
    fires when it first exceeds `limit` (quadrotor_dynamics.py:554-558); reproduce that accumulation literally."""
    acc, n = 0.0, 0
    while True:
        acc += dt
        n += 1
        if acc > limit:
            return n
