#!/usr/bin/env python
"""bench.py -- drone-steps/sec of the batched quadrotor-swarm step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                 # CPU arm: the oracle port on all host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU, weak scaling

Workload = the simulator leg of BASELINE.json configs[4] ("synthetic random-action throughput sweep: 65536 envs x
8 quads"), with the environment of configs[1]: 8-quad swarm, static_same_goal, pos_vel observations of the 6 nearest
neighbours (obs 54), sensor + thrust noise on, episodes of 1500 control steps (the timed window contains auto-resets),
synthetic i.i.d. U(-1,1) actions resident in HBM.  65536 envs PER GPU (weak scaling: env shards are independent, no
collective in the step).  A "step" is one control step of every env = one launch of the fused step kernel.  The
per-step working set (260 MB of state + observations) is larger than the 126 MB L2, so steps run back to back without
an L2 flush.  The 4096-envs-per-GPU point of configs[1] is measured too (per-step events, L2 flushed between steps)
and reported under "cfg2_4096".  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
ENVS_CFG2 = 4096
AGENTS = 8
# Algorithmic HBM bytes per drone-step for this workload (restated in DESIGN.md "Roofline"):
# state 120 B read + 120 B written, goal 12 R, tick/flags 4 R+W, action 16 R, obs 54*4 W, reward 4 + done 1 W
ALGO_BYTES_PER_DRONE_STEP = 497.0
WORKLOAD = ("cfg5 simulator leg: 65536 envs x 8 quads per GPU, env = cfg2 (static_same_goal, pos_vel x6 neighbours, "
            "obs 54, noise on, ep 1500 steps)")


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic_per_launch():
    """dram bytes per launch of the step kernel from the committed `ncu --set full` capture, if any."""
    p = os.path.join(ROOT, "profiles", "step_kernel_cfg2_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(envs, agents, budget_s, threads):
    """The CPU oracle (float64 C port of the reference's step, oracle/quadsim_oracle.c) on `threads` host threads,
    one slice of independent envs per thread, same workload config.  Returns (drone-steps/s, description)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import OracleBatch
    from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig
    cfg = QuadSimConfig(num_envs=envs, num_agents=agents, seed=0)
    b = OracleBatch(cfg, threads=threads)
    b.reset()
    rs = np.random.RandomState(0)
    acts = rs.uniform(-1, 1, (8, envs * agents, 4))
    for i in range(3):
        b.step(acts[i])
    t0 = time.perf_counter()
    b.step(acts[3]); b.step(acts[4])
    per = (time.perf_counter() - t0) / 2
    n = int(max(5, min(20000, budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    for i in range(n):
        b.step(acts[i % 8])
    dt = time.perf_counter() - t0
    return envs * agents * n / dt, f"{envs} envs x {agents} drones x {n} control steps, {threads} threads, {dt:.1f} s"


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (the reference is pure Python and cannot
    travel to the GPU box, so this is its C port, the oracle).  Rank 0 only."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    envs = 512
    budget = min(20.0, max(2.0, 0.02 * (args.steps + args.warmup)))
    v, sample = cpu_port_throughput(envs, AGENTS, budget, threads)
    line = {
        "impl": "reference", "metric": "drone-steps/sec", "value": v, "unit": "drone-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ENVS_PER_GPU * AGENTS / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": ENVS_PER_GPU, "agents": AGENTS, "obs_dim": 54,
                   "note": "CPU arm: C port of the reference step (oracle), all host threads, bounded sample; "
                           "ms_per_step is the time this arm would need for one 65536-env step"},
        "cpu_baseline": {"value": v, "unit": "drone-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "drone-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU, help="override for size sweeps (not the bench line)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-small", action="store_true", help="skip the secondary points (4096-env cfg2, cfg3 obstacles)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg3 = obstacle scenario as the timed workload (profiling aid; the bench line is cfg2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from quad_swarm_rl_stable_baselines3_b200.config import QuadSimConfig
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def cfg3_config(n_envs):
        # BASELINE.json configs[2]: 8x8 m obstacle area, density 0.2 -> 12 obstacles of 0.6 m, SDF obs, downwash, 2 neighbours, obs 40
        return QuadSimConfig(num_envs=n_envs, num_agents=AGENTS, quads_mode="mix", use_obstacles=True, use_downwash=True,
                             obs_repr="xyz_vxyz_R_omega_floor", neighbor_visible_num=2, seed=0, env_id_offset=rank * n_envs)

    def make(n_envs, workload="cfg2"):
        if workload == "cfg3":
            cfg = cfg3_config(n_envs)
        elif workload == "mix":      # the upstream training recipe --quads_mode=mix (swarm_rl/runs/quad_multi_mix_baseline.py:13-16)
            cfg = QuadSimConfig(num_envs=n_envs, num_agents=AGENTS, quads_mode="mix", seed=0, env_id_offset=rank * n_envs)
        else:
            cfg = QuadSimConfig(num_envs=n_envs, num_agents=AGENTS, seed=0, env_id_offset=rank * n_envs)
        sim = QuadSwarmSim(cfg, device=dev)
        sim.want_terminal_obs = False
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        pool = (torch.rand((POOL, n_envs * AGENTS, 4), device=dev, generator=gen) * 2.0 - 1.0).contiguous()
        sim.reset()
        return sim, pool

    POOL = 8
    n_envs = args.envs_per_gpu
    nd = n_envs * AGENTS
    sim, act_pool = make(n_envs, args.workload)

    # ---- device-resident path: value.  K back-to-back steps, one CUDA-event pair on the launching stream ---------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # nvidia-smi takes ~1 s to deliver its first sample: start it before the warm-up
    for i in range(args.warmup):
        sim.step(act_pool[i % POOL])
    barrier()
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        sim.step(act_pool[i % POOL])
    e1.record(stream)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    total_ms = e0.elapsed_time(e1)
    launches = sim.launch_count - l0
    clocks = None
    if rank == 0:
        if wall_s < 3.0:      # keep the GPU under the same load a little longer so that nvidia-smi samples it
            t_end = time.perf_counter() + 3.0
            j = 0
            while time.perf_counter() < t_end:
                sim.step(act_pool[j % POOL]); j += 1
            torch.cuda.synchronize(dev)
        clocks = sampler.stop()

    # ---- host-buffer path through the public API: e2e (pinned numpy in, pinned numpy out, every step) ------------
    host_acts = []
    for i in range(POOL):
        t = torch.empty((nd, 4), dtype=torch.float32, pin_memory=True)
        t.copy_(act_pool[i])
        host_acts.append(t.numpy())
    out_t = (torch.empty((nd, sim.D), dtype=torch.float32, pin_memory=True),
             torch.empty((nd,), dtype=torch.float32, pin_memory=True), torch.empty((nd,), dtype=torch.uint8, pin_memory=True))
    out = tuple(t.numpy() for t in out_t)
    e2e_steps = max(10, min(args.steps, 200))
    for i in range(3):
        sim.step_host(host_acts[i % POOL], out)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(e2e_steps):
        sim.step_host(host_acts[i % POOL], out)
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))     # the call is synchronous: host clock == device clock

    # ---- cfg2 point: 4096 envs per GPU, per-step events, L2 flushed between steps -------------------------------
    small_ms = None
    if not args.skip_small:
        del sim, act_pool
        sim2, pool2 = make(ENVS_CFG2)
        flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        for i in range(20):
            sim2.step(pool2[i % POOL])
        ns = min(args.steps, 300)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ns)]
        barrier()
        for i in range(ns):
            flush_buf.fill_(float(i))
            ev[i][0].record(stream)
            sim2.step(pool2[i % POOL])
            ev[i][1].record(stream)
        barrier()
        small_ms = sum(a.elapsed_time(b) for a, b in ev) / ns
        d_small = sim2.D
        del sim2, pool2, flush_buf
        # configs[2] (obstacle scenario), same 65536 envs per GPU, back-to-back steps
        sim3, pool3 = make(n_envs, "cfg3")
        for i in range(20):
            sim3.step(pool3[i % POOL])
        ns3 = min(args.steps, 300)
        barrier()
        e0.record(stream)
        for i in range(ns3):
            sim3.step(pool3[i % POOL])
        e1.record(stream)
        barrier()
        cfg3_ms = e0.elapsed_time(e1) / ns3
        del sim3, pool3
        # the formation-scenario kernel variant: quads_mode=mix, 9 scenarios drawn per episode
        simm, poolm = make(n_envs, "mix")
        for i in range(20):
            simm.step(poolm[i % POOL])
        barrier()
        e0.record(stream)
        for i in range(ns3):
            simm.step(poolm[i % POOL])
        e1.record(stream)
        barrier()
        mix_ms = e0.elapsed_time(e1) / ns3
        del simm, poolm
        # configs[0] family: the fork env sb_train.py trains on (4 chasers, PID pre-controller, 8 control steps per call)
        fcfg = QuadSimConfig.fork_default(num_envs=n_envs, seed=0, env_id_offset=rank * n_envs)
        simf = QuadSwarmSim(fcfg, device=dev)
        simf.want_terminal_obs = False
        poolf = (torch.rand((POOL, n_envs * fcfg.num_agents, 2), device=dev) * 2.0 - 1.0).contiguous()
        simf.reset()
        for i in range(10):
            simf.step(poolf[i % POOL])
        nsf = min(args.steps, 100)
        barrier()
        e0.record(stream)
        for i in range(nsf):
            simf.step(poolf[i % POOL])
        e1.record(stream)
        barrier()
        fork_ms = e0.elapsed_time(e1) / nsf
        fork_agents, fork_sub = fcfg.num_agents, fcfg.fork.substeps
        del simf, poolf
    else:
        d_small = 54
        cfg3_ms = fork_ms = mix_ms = 0.0
        fork_agents, fork_sub = 4, 8

    t = torch.tensor([total_ms, e2e_ms, small_ms or 0.0, cfg3_ms, fork_ms, mix_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, small_ms_max, cfg3_ms, fork_ms, mix_ms = (float(v) for v in t)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * nd / (ms_per_step * 1e-3)
        e2e_value = world * nd / (e2e_ms / e2e_steps * 1e-3)
        peak, peak_src = measured_hbm_peak()
        bytes_per_launch = ALGO_BYTES_PER_DRONE_STEP * nd
        achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
        cpu = None
        if not args.skip_cpu:
            threads = os.cpu_count() or 1
            v, sample = cpu_port_throughput(512, AGENTS, 12.0, threads)
            cpu = {"value": v, "unit": "drone-steps/s", "cores": threads, "kind": "port", "sample": sample}
        line = {
            "metric": "drone-steps/sec", "value": value, "unit": "drone-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n_envs, "agents": AGENTS, "obs_dim": d_small,
                       "act_dim": 4, "l2": "inputs larger than L2 (260 MB touched per step vs 126 MB L2), no flush",
                       "timing": "one CUDA-event pair around the K back-to-back steps on the launching stream, max over ranks",
                       "parallelism": f"env-sharded x{world}, no collective in the step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "drone-steps/s", "h2d_bytes_per_step": nd * 4 * 4,
                    "d2h_bytes_per_step": nd * (d_small * 4 + 4 + 1), "steps": e2e_steps,
                    "api": "QuadSwarmSim.step_host -> qs_step_host: pinned numpy actions H2D, step kernel, "
                           "obs/rew/done D2H into pinned numpy, stream sync, every step"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": committed_traffic_per_launch(), "algorithmic_bytes_per_launch": bytes_per_launch,
                         "kernel": "qs::step_kernel<8>", "peak_source": peak_src},
            "cpu_baseline": cpu,
        }
        if small_ms is not None:
            nd2 = ENVS_CFG2 * AGENTS
            ach2 = ALGO_BYTES_PER_DRONE_STEP * nd2 / (small_ms_max * 1e-3) / 1e9
            line["cfg2_4096"] = {"workload": "configs[1]: 4096 envs x 8 quads per GPU, same env", "ms_per_step": small_ms_max,
                                 "value": world * nd2 / (small_ms_max * 1e-3), "unit": "drone-steps/s",
                                 "l2": "flushed between timed steps (256 MiB write), per-step CUDA events",
                                 "roofline_frac": ach2 / peak,
                                 "note": "32768 threads = 0.43 waves of 148 SMs: bounded by one thread's latency chain, not by HBM"}
            line["cfg3_obstacles"] = {"workload": "configs[2]: 65536 envs x 8 quads per GPU, 12 obstacles, SDF obs, downwash, obs 40",
                                      "ms_per_step": cfg3_ms, "value": world * nd / (cfg3_ms * 1e-3), "unit": "drone-steps/s",
                                      "roofline_frac": 453.0 * nd / (cfg3_ms * 1e-3) / 1e9 / peak,
                                      "algorithmic_bytes_per_drone_step": 453.0}
            line["mix_scenarios"] = {"workload": "upstream training recipe quads_mode=mix: 65536 envs x 8 quads per GPU, one of 9 formation "
                                                 "scenarios per episode (goals moving every step in 3 of them), obs 54",
                                     "ms_per_step": mix_ms, "value": world * nd / (mix_ms * 1e-3), "unit": "drone-steps/s",
                                     "roofline_frac": (ALGO_BYTES_PER_DRONE_STEP + 16.0 + 12.0) * nd / (mix_ms * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_drone_step": ALGO_BYTES_PER_DRONE_STEP + 16.0 + 12.0,
                                     "note": "+16 B goal written every step, +96 B scenario row per env read (12 B per drone)"}
            line["fork_k4"] = {"workload": "fork env of sb_train.py: 65536 envs x 4 chasers per GPU, dynamic_repulsive, one call = 8 control steps",
                               "ms_per_call": fork_ms, "value": world * n_envs * fork_agents * fork_sub / (fork_ms * 1e-3),
                               "unit": "drone-steps/s (control steps)",
                               "agent_steps_per_s": world * n_envs * fork_agents / (fork_ms * 1e-3)}
        if args.workload != "cfg2":
            line["config"]["workload"] = "PROFILING AID, not the bench line: " + args.workload
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
