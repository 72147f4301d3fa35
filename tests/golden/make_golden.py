#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference from /root/reference.

Run in the build container (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Two kinds of fixtures:

  * ``trace_<name>.npz`` -- `QuadrotorEnvMulti.step` traces with numba's JIT disabled so that every random draw
    on the hot path is taped (oracle/ref_harness.py).  Per step: actions, the dynamics snapshot after the step,
    returned obs/reward/done, and the unit draws consumed (normals / uniforms / choice ids, reference order).
  * ``dyn_jit.npz`` -- `QuadrotorDynamics.step` driven directly with the numba JIT ON (the reference's real code
    path), thrust noise off: free flight, the 100-sub-step SVD re-orthonormalisation, floor contact and sliding.

tests/test_oracle_golden.py replays them through oracle/quadsim_oracle.c.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

STATE_KEYS = ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "ou", "on_floor", "crashed_floor",
              "crashed_wall", "crashed_ceiling", "goal")

# name -> (env kwargs, steps, action scale/bias recipe)
TRACES = {
    # cfg2 shape: 8 quads, static_same_goal, 6 nearest neighbours, noise on; short episodes so resets are covered
    "cfg2_k8": dict(env=dict(num_agents=8, ep_time=1.2), steps=260, act="uniform"),
    # tight room: drones spawn outside the box -> wall / ceiling bounces, floor, pair collisions
    "smallroom_k8": dict(env=dict(num_agents=8, room_dims=(3.0, 3.0, 3.0), ep_time=1.0), steps=220, act="high"),
    # crowded: 16 quads in a small spawn volume -> drone-drone impulses (incl. drones in two new pairs)
    "crowd_k16": dict(env=dict(num_agents=16, room_dims=(4.0, 4.0, 6.0), ep_time=1.0, neighbor_visible_num=6),
                      steps=150, act="hover"),
    # cfg3 shape: obstacles (mix -> o_random | o_static_same_goal), SDF obs, downwash, 2 neighbours
    "cfg3_obst_k8": dict(env=dict(num_agents=8, quads_mode="mix", use_obstacles=True, use_downwash=True,
                                  obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, ep_time=1.0,
                                  rew_coeff=dict(pos=1.0, effort=0.05, spin=0.1, vel=0.0, crash=1.0, orient=1.0,
                                                 yaw=0.0, quadcol_bin=5.0, quadcol_bin_smooth_max=4.0,
                                                 quadcol_bin_obst=5.0)),
                         steps=330, act="hover"),
    # cfg4 shape: 32 quads
    "cfg4_k32": dict(env=dict(num_agents=32, ep_time=0.5), steps=70, act="uniform"),
    # no sensor noise / all neighbours visible / wall obs
    "nonoise_k4": dict(env=dict(num_agents=4, sense_noise=None, neighbor_visible_num=-1,
                                obs_repr="xyz_vxyz_R_omega_wall", ep_time=0.6), steps=130, act="uniform"),
}


def gen_traces():
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    import numpy as np
    import ref_harness as rh

    tape = rh.Tape(seed=1234)
    rh.install_tape(tape)
    for name, spec in TRACES.items():
        K = spec["env"]["num_agents"]
        env = rh.make_upstream_env(tape=tape, **spec["env"])
        rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        rec = {k: [] for k in ("actions", "obs", "rew", "done", "tn", "tu", "tc", "n_tn", "n_tu", "n_tc", "tick",
                               "obst_xy", "scenario", "ep_stats", "ep_step")}
        snaps = {k: [] for k in STATE_KEYS}

        def push_tape(m):
            k, v = tape.since(m)
            for key, kind in (("tn", 0), ("tu", 1), ("tc", 2)):
                vals = v[k == kind]
                rec[key].append(vals)
                rec["n_" + key].append(len(vals))

        def push_snap():
            s = rh.snapshot(env)
            for k in STATE_KEYS:
                snaps[k].append(s[k])
            rec["tick"].append(int(s["tick"][0]))
            if env.use_obstacles:
                rec["obst_xy"].append(np.array(env.obstacles.pos_arr)[:, :2].copy())
                rec["scenario"].append(env.scenario.scenario.__class__.__name__)

        m = tape.mark()
        obs, _ = env.reset()
        push_tape(m)
        push_snap()
        rec["obs"].append(np.array(obs, dtype=np.float64))
        events = dict(done=0, impulse=0)
        for s in range(spec["steps"]):
            if spec["act"] == "uniform":
                a = rs.uniform(-1.0, 1.0, (K, 4))
            elif spec["act"] == "high":          # mostly climbing -> ceiling / wall hits
                a = rs.uniform(-0.2, 1.3, (K, 4))
            else:                                # near hover, small differential -> long free flight, collisions
                a = 0.05 + rs.uniform(-0.15, 0.15, (K, 4))
            m = tape.mark()
            obs, rew, done, infos = env.step(a)
            push_tape(m)
            push_snap()
            rec["actions"].append(a)
            rec["obs"].append(np.array(obs, dtype=np.float64))
            rec["rew"].append(np.array(rew, dtype=np.float64))
            rec["done"].append(np.array(done, dtype=bool))
            events["done"] += int(any(done))
            if any(done):
                # infos[i]['episode_extra_stats'] of the finished episode (quadrotor_multi.py:739-831), summed over agents where
                # the value is per agent -> the aggregate qs_stats keeps
                es = [inf["episode_extra_stats"] for inf in infos]
                e0 = es[0]
                rec["ep_step"].append(s)
                rec["ep_stats"].append([
                    e0["num_collisions"], e0["num_collisions_after_settle"], e0["num_collisions_final_5_s"],
                    e0["num_collisions_with_room"], e0["num_collisions_with_floor"], e0["num_collisions_with_wall"],
                    e0["num_collisions_with_ceiling"], e0.get("num_collisions_obst_quad", 0),
                    e0.get("num_collisions_obst_quad_after_settle", 0),
                    round(e0["metric/agent_success_rate"] * K), round(e0["metric/agent_deadlock_rate"] * K),
                    round(e0["metric/agent_col_rate"] * K),
                    sum(x["distance_to_goal_1s"] for x in es), sum(x["distance_to_goal_3s"] for x in es),
                    sum(x["distance_to_goal_5s"] for x in es)])
        out = dict(
            actions=np.array(rec["actions"]), obs=np.array(rec["obs"]), rew=np.array(rec["rew"]),
            done=np.array(rec["done"]), tick=np.array(rec["tick"]),
            tn=np.concatenate(rec["tn"]), tu=np.concatenate(rec["tu"]), tc=np.concatenate(rec["tc"]),
            n_tn=np.array(rec["n_tn"]), n_tu=np.array(rec["n_tu"]), n_tc=np.array(rec["n_tc"]),
            ep_step=np.array(rec["ep_step"], dtype=np.int64), ep_stats=np.array(rec["ep_stats"], dtype=np.float64),
        )
        for k in STATE_KEYS:
            out["s_" + k] = np.array(snaps[k])
        if env.use_obstacles:
            out["obst_xy"] = np.array(rec["obst_xy"])
            out["scenario"] = np.array(rec["scenario"])
        out["env_kwargs"] = np.array(repr(spec["env"]))
        path = os.path.join(HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: steps={spec['steps']} dones={events['done']} floor={int(out['s_on_floor'].any(axis=1).sum())} "
              f"wall={int(out['s_crashed_wall'].sum())} ceil={int(out['s_crashed_ceiling'].sum())} "
              f"draws N={len(out['tn'])} U={len(out['tu'])} C={len(out['tc'])} -> {os.path.getsize(path) // 1024} KiB")


# fork env (quadrotor_multi_rewards.py) traces: name -> (QuadrotorEnvConfig overrides, VecEnv steps, capture-radius schedule)
FORK_TRACES = {
    # sb_train.py defaults: 4 chasers, dist_angle neighbours, capture radius 3.0 -> frequent immediate captures + timeouts
    "fork_k4": dict(env=dict(num_agents=4, episode_duration=1.6), steps=110, radius={40: 1.0, 80: 2.6}),
    # BASELINE configs[0]: single quadrotor
    "fork_k1": dict(env=dict(num_agents=1, episode_duration=1.0, initial_capture_radius=2.4), steps=60, radius={}),
    # sangle reprs, 3 nearest of 7 neighbours
    "fork_k8_sangle": dict(env=dict(num_agents=8, episode_duration=0.8, initial_capture_radius=2.2,
                                    obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot",
                                    neighbor_obs_type="dist_sangle", neighbor_visible_num=3), steps=50, radius={}),
    # the author's current sweep (sb_train.py:122-137): camera-model neighbour observations, here with pixel noise on
    "fork_k4_cam": dict(env=dict(num_agents=4, episode_duration=0.8, initial_capture_radius=2.4,
                                 obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot", neighbor_obs_type="ndist_nsangle"),
                        steps=40, radius={}),
    # camera model + ranking (2 nearest of 5): the ranking pass draws its own pixel noise
    "fork_k6_cam_v2": dict(env=dict(num_agents=6, episode_duration=0.8, initial_capture_radius=2.2, pixel_noise_cam=1.0,
                                    n_cameras=4, neighbor_obs_type="ndist_nsangle", neighbor_visible_num=2), steps=30, radius={}),
    # relative-heading neighbour types
    "fork_k4_heading": dict(env=dict(num_agents=4, episode_duration=0.8, initial_capture_radius=2.4,
                                     neighbor_obs_type="dist_angle_heading"), steps=40, radius={}),
    "fork_k5_sheading_v3": dict(env=dict(num_agents=5, episode_duration=0.8, initial_capture_radius=2.2,
                                         obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot",
                                         neighbor_obs_type="dist_sangle_sheading", neighbor_visible_num=3), steps=30, radius={}),
}
FORK_STATE_KEYS = STATE_KEYS + ("pid", "angle", "ang_vel", "evader", "heading")


def gen_fork_traces():
    """The env sb_train.py trains on, driven the way SubprocVecEnvCustom's worker drives it
    (swarm_rl/env_wrappers/subproc_vec_env_custom.py:35-52): step, and on any(done) keep the terminal observation and reset."""
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    import io
    import contextlib
    import numpy as np
    import ref_harness as rh

    tape = rh.Tape(seed=4321)
    rh.install_tape(tape)
    for name, spec in FORK_TRACES.items():
        K = spec["env"]["num_agents"]
        with contextlib.redirect_stdout(io.StringIO()):
            env = rh.make_fork_env(tape=tape, **spec["env"])
        rs = np.random.RandomState(sum(map(ord, name)))
        rec = {k: [] for k in ("actions", "obs", "rew", "done", "term", "success", "radius", "tn", "tu", "n_tn", "n_tu", "tick")}
        snaps = {k: [] for k in FORK_STATE_KEYS}

        def push_tape(m):
            k, v = tape.since(m)
            assert not (k == 2).any()
            for key, kind in (("tn", 0), ("tu", 1)):
                vals = v[k == kind]
                rec[key].append(vals)
                rec["n_" + key].append(len(vals))

        def push_snap():
            sn = rh.fork_snapshot(env)
            for k in FORK_STATE_KEYS:
                snaps[k].append(sn[k])
            rec["tick"].append(int(sn["tick"][0]))

        m = tape.mark()
        with contextlib.redirect_stdout(io.StringIO()):
            obs, _ = env.reset()
        push_tape(m)
        push_snap()
        rec["obs"].append(np.array(obs, dtype=np.float64))
        n_done = n_succ = 0
        for s in range(spec["steps"]):
            if s in spec["radius"]:
                env.set_capture_radius(spec["radius"][s])
            rec["radius"].append(float(env.capture_radius))
            a = rs.uniform(-1.0, 1.0, (K, 2))
            m = tape.mark()
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                obs, rew, done, infos = env.step(a)
                term = np.array(obs, dtype=np.float64)
                succ = -1
                if any(done):
                    obs, info = env.reset()
                    succ = int(info["success"])
                    n_done += 1
                    n_succ += succ
            push_tape(m)
            push_snap()
            rec["actions"].append(a)
            rec["obs"].append(np.array(obs, dtype=np.float64))
            rec["rew"].append(np.array(rew, dtype=np.float64))
            rec["done"].append(np.array(done, dtype=bool))
            rec["term"].append(term)
            rec["success"].append(succ)
        out = dict(actions=np.array(rec["actions"]), obs=np.array(rec["obs"]), rew=np.array(rec["rew"]),
                   done=np.array(rec["done"]), term=np.array(rec["term"]), success=np.array(rec["success"]),
                   radius=np.array(rec["radius"]), tick=np.array(rec["tick"]),
                   tn=np.concatenate(rec["tn"]), tu=np.concatenate(rec["tu"]),
                   n_tn=np.array(rec["n_tn"]), n_tu=np.array(rec["n_tu"]))
        for k in FORK_STATE_KEYS:
            out["s_" + k] = np.array(snaps[k])
        out["env_kwargs"] = np.array(repr(spec["env"]))
        path = os.path.join(HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: steps={spec['steps']} dones={n_done} captures={n_succ} draws N={len(out['tn'])} U={len(out['tu'])} "
              f"-> {os.path.getsize(path) // 1024} KiB")


def gen_dyn_jit():
    """JIT ON: QuadrotorDynamics.step (quadrotor_dynamics.py:215-221) on one drone, thrust noise off."""
    import numpy as np
    import ref_harness as rh
    rh.install_stubs()
    from gym_art.quadrotor_multi.quadrotor_dynamics import QuadrotorDynamics
    from gym_art.quadrotor_multi.quad_utils import rpy2R
    import gym_art.quadrotor_multi.quadrotor_randomization as qr

    params = qr.Crazyflie().sample()
    params["noise"]["thrust_noise_ratio"] = 0.0
    room_box = np.array([[-5.0, -5.0, 0.0], [5.0, 5.0, 10.0]])
    rs = np.random.RandomState(7)
    runs = []
    for run in range(4):
        dyn = QuadrotorDynamics(model_params=params, dynamics_steps_num=2, room_box=room_box, use_numba=True, dt=0.005)
        pos = np.array([rs.uniform(-1, 1), rs.uniform(-1, 1), rs.uniform(0.3, 2.5)])
        vel = rs.uniform(-0.5, 0.5, 3)
        rot = rpy2R(*rs.uniform(-0.4, 0.4, 3))
        if run == 3:                                   # upside down -> random-yaw branch never fires w/o RNG tap; keep z>0
            rot = rpy2R(0.2, 2.6, -1.0)
            pos[2] = 3.0
        omega = rs.uniform(-1.0, 1.0, 3)
        dyn.set_state(pos, vel, rot, omega)
        dyn.reset()
        dyn.on_floor = False
        steps = 260
        cmds = rs.uniform(0.0, 1.0, (steps, 4)) * (0.6 if run % 2 == 0 else 1.0)
        tr = {k: [] for k in ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "on_floor", "crashed_floor", "acc")}

        def snap():
            tr["pos"].append(np.array(dyn.pos, dtype=np.float64)); tr["vel"].append(np.array(dyn.vel, dtype=np.float64))
            tr["rot"].append(np.array(dyn.rot, dtype=np.float64).reshape(9)); tr["omega"].append(np.array(dyn.omega, dtype=np.float64))
            tr["rot_damp"].append(np.array(dyn.thrust_rot_damp, dtype=np.float64))
            tr["cmds_damp"].append(np.array(dyn.thrust_cmds_damp, dtype=np.float64))
            tr["on_floor"].append(bool(dyn.on_floor)); tr["crashed_floor"].append(bool(dyn.crashed_floor))
            tr["acc"].append(np.array(dyn.acc, dtype=np.float64))

        snap()
        for s in range(steps):
            dyn.step(cmds[s], 0.005)
            snap()
        runs.append(dict(cmds=cmds, **{k: np.array(v) for k, v in tr.items()}))
    out = {}
    for r, d in enumerate(runs):
        for k, v in d.items():
            out[f"r{r}_{k}"] = v
    out["n_runs"] = np.array(len(runs))
    out["consts"] = np.array([dyn.mass, *dyn.inertia, *dyn.thrust_max, *dyn.torque_max, dyn.arm, dyn.motor_tau_up,
                              *dyn.prop_crossproducts.reshape(-1)])
    path = os.path.join(HERE, "dyn_jit.npz")
    np.savez_compressed(path, **out)
    print("dyn_jit:", {r: int(d["on_floor"].sum()) for r, d in enumerate(runs)}, "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "traces":
        gen_traces()
    elif which == "fork":
        gen_fork_traces()
    elif which == "dyn":
        gen_dyn_jit()
    else:   # separate interpreters: NUMBA_DISABLE_JIT must be decided before numba is imported
        subprocess.check_call([sys.executable, __file__, "dyn"])
        subprocess.check_call([sys.executable, __file__, "traces"])
        subprocess.check_call([sys.executable, __file__, "fork"])
