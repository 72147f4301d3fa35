"""B200-native batched quadrotor-swarm simulator (drop-in for the hot path of priban42/quad-swarm-rl-stable-baselines3).

    from quad_swarm_rl_stable_baselines3_b200 import QuadSimConfig, QuadSwarmSim, QuadSwarmVecEnv
"""
from .config import QuadSimConfig  # noqa: F401


def __getattr__(name):          # torch is only imported when the simulator classes are asked for
    if name == "QuadSwarmSim":
        from .sim import QuadSwarmSim
        return QuadSwarmSim
    if name == "QuadSwarmVecEnv":
        from .vec_env import QuadSwarmVecEnv
        return QuadSwarmVecEnv
    raise AttributeError(name)
