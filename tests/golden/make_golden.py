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
    # single drone with obstacles: mix draws from the one-entry list ['o_random'] whatever the mode index
    "obst_k1": dict(env=dict(num_agents=1, quads_mode="mix", use_obstacles=True, obs_repr="xyz_vxyz_R_omega_floor",
                             neighbor_visible_num=0, neighbor_obs_type="none", ep_time=0.3), steps=200, act="hover"),
    # no sensor noise / all neighbours visible / wall obs
    "nonoise_k4": dict(env=dict(num_agents=4, sense_noise=None, neighbor_visible_num=-1,
                                obs_repr="xyz_vxyz_R_omega_wall", ep_time=0.6), steps=130, act="uniform"),
}


# Formation scenarios.  Sensor noise off and float32 observations keep the files small: the state, goal and
# reward traces stay float64 and pin the arithmetic; the scenario object's own state is recorded next to them.
_SC = dict(sense_noise=None, neighbor_visible_num=2)
SCENARIO_TRACES = {
    # every formation over many short episodes: circle / sphere / grid / cube goal layouts, spawn at the shuffled goals
    "scen_static_diff_k8": dict(env=dict(num_agents=8, quads_mode="static_diff_goal", ep_time=0.12, **_SC), steps=420, act="hover"),
    # 12 drones: two circle layers, 3x4 grid, cube with floor dimension 2
    "scen_static_diff_k12": dict(env=dict(num_agents=12, quads_mode="static_diff_goal", ep_time=0.08, **_SC), steps=150, act="hover"),
    # goal teleports every 4-6 s
    "scen_dyn_same_k3": dict(env=dict(num_agents=3, quads_mode="dynamic_same_goal", ep_time=6.3, **_SC), steps=680, act="hover"),
    "scen_dyn_diff_k4": dict(env=dict(num_agents=4, quads_mode="dynamic_diff_goal", ep_time=6.3, **_SC), steps=680, act="hover"),
    "scen_swap_k3": dict(env=dict(num_agents=3, quads_mode="swap_goals", ep_time=6.3, **_SC), steps=680, act="hover"),
    "scen_runaway_k5": dict(env=dict(num_agents=5, quads_mode="run_away", ep_time=4.3, **_SC), steps=900, act="hover"),
    "scen_swarm_k6": dict(env=dict(num_agents=6, quads_mode="swarm_vs_swarm", ep_time=6.3, **_SC), steps=680, act="hover"),
    # halves of 2 drones: a sphere of fewer than 3 drones still yields 3 goal rows (scenarios/utils.py:77-80)
    "scen_swarm_k4": dict(env=dict(num_agents=4, quads_mode="swarm_vs_swarm", ep_time=0.08, **_SC), steps=260, act="hover"),
    # formation size breathing every step, direction flips at +-highest_formation_size
    "scen_dynform_k5": dict(env=dict(num_agents=5, quads_mode="dynamic_formations", ep_time=4.0, **_SC), steps=900, act="hover"),
    "scen_lissajous_k3": dict(env=dict(num_agents=3, quads_mode="ep_lissajous3D", ep_time=1.0, **_SC), steps=230, act="hover"),
    # quadratic Bezier arcs resampled at tick 1 and every 5 s (rejection loop on the room bounds)
    "scen_bezier_k3": dict(env=dict(num_agents=3, quads_mode="ep_rand_bezier", ep_time=5.4, **_SC), steps=600, act="hover"),
    # the upstream training recipe --quads_mode=mix: a fresh scenario object per episode
    "scen_mix_k4": dict(env=dict(num_agents=4, quads_mode="mix", ep_time=0.25, **_SC), steps=1000, act="hover"),
    "scen_mix_k1": dict(env=dict(num_agents=1, quads_mode="mix", ep_time=0.2, sense_noise=None, neighbor_visible_num=0,
                                 neighbor_obs_type="none"), steps=400, act="hover"),
}
FORMATIONS = ("circle_horizontal", "circle_vertical_xz", "circle_vertical_yz", "sphere", "grid_horizontal",
              "grid_vertical_xz", "grid_vertical_yz", "cube")


def scenario_row(env):
    """The live scenario object's state in the QS_SC_* layout of include/quadsim.h (+ its class name)."""
    import numpy as np
    sc = env.scenario.scenario if hasattr(env.scenario, "scenario") and env.scenario.scenario is not None else env.scenario
    row = np.zeros(24)
    row[1] = FORMATIONS.index(sc.formation)
    row[2], row[3], row[4], row[5] = sc.formation_size, sc.layer_dist, sc.highest_formation_size, sc.lowest_formation_size
    row[6:9] = sc.formation_center
    row[9] = getattr(sc, "control_step_for_sec", 0)
    row[10] = float(getattr(sc, "increase_formation_size", 0))
    row[11] = getattr(sc, "control_speed", 0.0)
    if getattr(sc, "goal_center_1", None) is not None:
        row[12:15], row[15:18] = sc.goal_center_1, sc.goal_center_2
    return sc.__class__.__name__[len("Scenario_"):], row


def gen_formations():
    """QuadrotorScenario.generate_goals of the reference for every formation x swarm size (scenarios/base.py:42-116)."""
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    import numpy as np
    import ref_harness as rh
    rh.install_stubs()
    from gym_art.quadrotor_multi.scenarios.base import QuadrotorScenario
    out, index = {}, []
    rs = np.random.RandomState(5)
    for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 17, 24, 27, 32):
        for fi, form in enumerate(FORMATIONS):
            sc = QuadrotorScenario("static_diff_goal", envs=[], num_agents=n, room_dims=(10, 10, 10), rng=np.random.default_rng(0))
            sc.formation = form
            sc.num_agents_per_layer = 50 if form.startswith("grid") else 8
            sc.formation_size = float(rs.uniform(-0.5, 1.5))
            center, layer = rs.uniform(-2, 2, 3) + np.array([0, 0, 3.0]), float(rs.uniform(0.1, 0.6))
            goals = np.array(sc.generate_goals(num_agents=n, formation_center=center, layer_dist=layer), dtype=np.float64)
            index.append([n, fi, sc.formation_size, layer, *center, goals.shape[0]])
            out[f"g{len(index) - 1}"] = goals
    out["index"] = np.array(index)
    path = os.path.join(HERE, "formations.npz")
    np.savez_compressed(path, **out)
    print("formations:", len(index), "cases ->", os.path.getsize(path) // 1024, "KiB")


def trace_seed(name, base):
    """Every trace draws from its own stream (seeded from its name), so adding, removing or reordering traces never changes the
    others: `python make_golden.py traces <name>` reproduces the committed file of that name alone."""
    import zlib
    return (zlib.crc32(name.encode()) ^ base) & 0x7FFFFFFF


def reseed(tape, name, base, salt=0):
    """salt: a per-trace constant in the trace table, chosen so that the trace covers the events its test asserts (a capture)."""
    import numpy as np
    tape.rs = np.random.RandomState(trace_seed(name if not salt else f"{name}#{salt}", base))
    tape.kinds.clear()
    tape.vals.clear()


def gen_traces(traces=None, tape_seed=1234, compact=False, out_dir=None):
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    import numpy as np
    import ref_harness as rh

    tape = rh.Tape(seed=tape_seed)
    rh.install_tape(tape)
    for name, spec in (TRACES if traces is None else traces).items():
        K = spec["env"]["num_agents"]
        reseed(tape, name, tape_seed)
        env = rh.make_upstream_env(tape=tape, **spec["env"])
        rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        rec = {k: [] for k in ("actions", "obs", "rew", "done", "tn", "tu", "tc", "n_tn", "n_tu", "n_tc", "tick",
                               "obst_xy", "scenario", "ep_stats", "ep_step", "sc_row", "ep_detail", "ep_scenario")}
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
            elif compact:
                nm, row = scenario_row(env)
                rec["scenario"].append(nm)
                rec["sc_row"].append(row)

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
            if compact:
                a = a.astype(np.float32).astype(np.float64)
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
                if compact:
                    # the whole record per agent: distances, the two extra rates and the scenario the keys are prefixed with
                    sname = env.scenario.name()[9:] if False else [k for k in e0 if k.endswith("/agent_success_rate") and not k.startswith("metric/")][0].split("/")[0]
                    rec["ep_detail"].append(np.array([[x["distance_to_goal_1s"], x["distance_to_goal_3s"], x["distance_to_goal_5s"],
                                                       x["metric/agent_neighbor_col_rate"], x["metric/agent_obst_col_rate"],
                                                       x[f"{sname}/distance_to_goal_3s"], x[f"{sname}/num_collisions"]] for x in es]))
                    rec["ep_scenario"].append(sname)
                    rec["ep_keys"] = sorted(k.replace(sname, "<scenario>") for k in e0)
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
        if compact:
            out["scenario"] = np.array(rec["scenario"])
            out["sc_row"] = np.array(rec["sc_row"])
            out["ep_detail"] = np.array(rec["ep_detail"])
            out["ep_scenario"] = np.array(rec["ep_scenario"])
            out["ep_keys"] = np.array(rec["ep_keys"])
            out["obs"] = out["obs"].astype(np.float32)
            out["actions"] = out["actions"].astype(np.float32)
            for k in ("on_floor", "crashed_floor", "crashed_wall", "crashed_ceiling"):
                out["s_" + k] = np.packbits(out["s_" + k], axis=None)
            out["flag_shape"] = np.array(np.array(snaps["on_floor"]).shape)
            goal_moves = int((np.abs(np.diff(out["s_goal"], axis=0)).max(axis=(1, 2)) > 0).sum())
            print(f"    scenarios={sorted(set(rec['scenario']))} formations={sorted(set(int(r[1]) for r in rec['sc_row']))} "
                  f"steps with goal changes={goal_moves}")
        out["env_kwargs"] = np.array(repr(spec["env"]))
        path = os.path.join(out_dir or HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: steps={spec['steps']} dones={events['done']} floor={int(np.array(snaps['on_floor']).any(axis=1).sum())} "
              f"wall={int(np.array(snaps['crashed_wall']).sum())} ceil={int(np.array(snaps['crashed_ceiling']).sum())} "
              f"draws N={len(out['tn'])} U={len(out['tu'])} C={len(out['tc'])} -> {os.path.getsize(path) // 1024} KiB")


# fork env (quadrotor_multi_rewards.py) traces: name -> (QuadrotorEnvConfig overrides, VecEnv steps, capture-radius schedule)
FORK_TRACES = {
    # sb_train.py defaults: 4 chasers, dist_angle neighbours, capture radius 3.0 -> frequent immediate captures + timeouts
    "fork_k4": dict(env=dict(num_agents=4, episode_duration=1.6), steps=110, radius={40: 1.0, 80: 2.6}),
    # BASELINE configs[0]: single quadrotor
    "fork_k1": dict(salt=3, env=dict(num_agents=1, episode_duration=1.0, initial_capture_radius=2.4), steps=60, radius={}),
    # sangle reprs, 3 nearest of 7 neighbours
    "fork_k8_sangle": dict(env=dict(num_agents=8, episode_duration=0.8, initial_capture_radius=2.2,
                                    obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot",
                                    neighbor_obs_type="dist_sangle", neighbor_visible_num=3), steps=50, radius={}),
    # the author's current sweep (sb_train.py:122-137): camera-model neighbour observations, here with pixel noise on
    "fork_k4_cam": dict(salt=3, env=dict(num_agents=4, episode_duration=0.8, initial_capture_radius=2.4,
                                 obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot", neighbor_obs_type="ndist_nsangle"),
                        steps=40, radius={}),
    # camera model + ranking (2 nearest of 5): the ranking pass draws its own pixel noise
    "fork_k6_cam_v2": dict(env=dict(num_agents=6, episode_duration=0.8, initial_capture_radius=2.2, pixel_noise_cam=1.0,
                                    n_cameras=4, neighbor_obs_type="ndist_nsangle", neighbor_visible_num=2), steps=30, radius={}),
    # the noisy self representation (get_state.py:190-224; commented out in sb_train.py:123 but a valid cfg.obs_repr), pixel noise on
    "fork_k3_nself": dict(salt=5, env=dict(num_agents=3, episode_duration=0.8, initial_capture_radius=2.3, pixel_noise_cam=0.7,
                                   obs_repr="cdist_cdistdot_ndist_distdot_nsangle_angledot", neighbor_obs_type="ndist_nsangle"),
                          steps=40, radius={}),
    # relative-heading neighbour types
    "fork_k4_heading": dict(env=dict(num_agents=4, episode_duration=0.8, initial_capture_radius=2.4,
                                     neighbor_obs_type="dist_angle_heading"), steps=40, radius={}),
    "fork_k5_sheading_v3": dict(env=dict(num_agents=5, episode_duration=0.8, initial_capture_radius=2.2,
                                         obs_repr="cdist_cdistdot_dist_distdot_sangle_angledot",
                                         neighbor_obs_type="dist_sangle_sheading", neighbor_visible_num=3), steps=30, radius={}),
}
FORK_STATE_KEYS = STATE_KEYS + ("pid", "angle", "ang_vel", "evader", "heading")


def gen_fork_traces(only=(), out_dir=None):
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
        if only and name not in only:
            continue
        K = spec["env"]["num_agents"]
        reseed(tape, name, 4321, spec.get("salt", 0))
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
        path = os.path.join(out_dir or HERE, f"trace_{name}.npz")
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
    out_dir = os.environ.get("QS_GOLDEN_OUT") or None       # tests regenerate into a scratch directory and compare
    if which == "traces":
        only = sys.argv[2:]
        gen_traces({k: v for k, v in TRACES.items() if not only or k in only}, out_dir=out_dir)
    elif which == "scenarios":
        only = sys.argv[2:]
        if not only:
            gen_formations()
        gen_traces({k: v for k, v in SCENARIO_TRACES.items() if not only or k in only}, tape_seed=2468, compact=True, out_dir=out_dir)
    elif which == "fork":
        gen_fork_traces(sys.argv[2:], out_dir=out_dir)
    elif which == "dyn":
        gen_dyn_jit()
    else:   # separate interpreters: NUMBA_DISABLE_JIT must be decided before numba is imported
        subprocess.check_call([sys.executable, __file__, "dyn"])
        subprocess.check_call([sys.executable, __file__, "traces"])
        subprocess.check_call([sys.executable, __file__, "fork"])
        subprocess.check_call([sys.executable, __file__, "scenarios"])
