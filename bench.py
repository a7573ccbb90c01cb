#!/usr/bin/env python
"""bench.py — agent-prediction-steps/s of the multi-agent predictive rollout on B200.

A "step" of the bench is one control tick (planCallback, panda_bimanual_control.cpp:329-369) of
the closed dry-run loop: evaluate the finished rollouts -> best agent -> move the real agent ->
re-seed every agent -> roll every agent out over the horizon. The metric counts EXECUTED
integration steps (agents x (horizon-1) on the synthetic workloads) per second.

  value   device-resident: pmaf_tick (one fused chain per tick), obstacles already in HBM,
          timed with CUDA events on the planner's stream (pmaf_timer_*), L2 flushed between ticks.
  e2e     the five reference-facing CfManager calls per tick through the C ABI with HOST buffers
          (obstacle lists re-uploaded every call, results read back), driven by the library's C++
          host loop pmaf_dry_run (the reference's caller is a C++ node), wall clock.
  workloads   the headline line is BASELINE.json configs[1] (C2); the same JSON line carries
          time-boxed sub-records of the other BASELINE configs — one GPU: C3, C5 and one GPU's
          share of C4; N GPUs: C4 with 8192 agents per GPU (N = 8 is C4 proper, 65 536 agents).
  sharded_parity   N > 1: before anything is timed, three ticks of a golden case run sharded over
          the N ranks and are compared bit for bit with the reference's golden vectors.
  --impl reference   the reference's own CPU implementation (oracle/_ref when built, else the
          C port) on all host threads, same workload and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import pmaf_b200  # noqa: E402,F401
from pmaf_b200 import loop, scenarios  # noqa: E402

METRIC = "agent-prediction-steps/sec"
WORKLOADS = {"c2": scenarios.c2, "c3": scenarios.c3, "c4": scenarios.c4, "c5": scenarios.c5,
             "c4s": lambda: scenarios.c4(8192)}  # c4s: one GPU's share of C4 (65536 agents over 8 GPUs)
SUB_TICKS = 10  # timed ticks of every sub-record (plus 3 warm-up ticks)


def host_threads():
    """Host threads this process may run on — NOT OpenMP's default, which torchrun pins to 1 (OMP_NUM_THREADS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(sc, n_gpus, extra=None):
    which = {"c2": "configs[1]", "c3": "configs[2]", "c4": "configs[3]", "c5": "configs[4]"}.get(sc.name[:2], "")
    cfg = {"workload": f"{sc.name}: {sc.num_agents} agents x {sc.num_obstacles} obstacles x horizon "
                       f"{sc.max_prediction_steps} (BASELINE.json {which}), closed dry-run loop, "
                       f"{'moving' if sc.feed_obstacles else 'static'} obstacles",
           "agents": sc.num_agents, "obstacles": sc.num_obstacles, "horizon": sc.max_prediction_steps,
           "agents_per_gpu": sc.num_agents // max(n_gpus, 1),
           "l2": "flushed between ticks (256 MiB memset on the planner stream, outside the timed region)"}
    cfg.update(extra or {})
    return cfg


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed regions (B200_PROFILING.md's clocks line).
    In-process NVML from a thread when `pynvml` is importable — a query costs microseconds; an `nvidia-smi -lms` child
    process re-initialises NVML state on every sample and was measured to stall this process's CUDA calls for
    milliseconds a few times per second — else the nvidia-smi loop of the recipe."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.025

    def __init__(self, gpu_index=0):
        self.sm, self.mx, self.reasons, self.how = [], [], set(), None
        self.p = self.f = self.thread = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self._stop = False

            def loop_():
                while not self._stop:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(self.PERIOD_S)

            import threading

            self.thread = threading.Thread(target=loop_, daemon=True)
            self.thread.start()
            self.how = f"in-process NVML, {int(1e3 * self.PERIOD_S)} ms period"
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
            self.how = "nvidia-smi -lms 100"
        except OSError:
            pass

    def wait_ready(self, timeout=8.0):
        """nvidia-smi's start-up (driver / NVML initialisation) can stall CUDA calls of other processes for
        milliseconds: wait for the first sample before anything is timed."""
        t0 = time.time()
        while time.time() - t0 < timeout:
            if (self.thread is not None and self.sm) or (self.p is not None and os.path.getsize(self.f.name) > 0):
                break
            if self.thread is None and self.p is None:
                break
            time.sleep(0.02)
        return self

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join()
        elif self.p is not None:
            time.sleep(0.1)
            self.p.terminate()
            self.p.wait()
            self.f.flush()
            rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
            os.unlink(self.f.name)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for r in rows:
                try:
                    self.sm.append(float(r[1])), self.mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm), "reasons": sorted(self.reasons),
                "how": self.how, "span": "all timed regions of this run (headline, roofline pass, e2e and sub-records)"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def measured_traffic(workload):
    """dram__bytes_read + dram__bytes_write of the rollout kernel, per launch, from the newest committed
    `ncu --set full` summary of the same workload (profiles/r*_ncu_rollout_<workload>_summary.json), or None."""
    import glob

    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_rollout_{workload}_summary.json")))
    try:
        return json.load(open(paths[-1])).get("dram_bytes_per_launch") if paths else None
    except (OSError, ValueError):
        return None


def cpu_planner(threads):
    from oracle import cpu_planners

    if cpu_planners.have_ref():
        return cpu_planners.RefPlanner(threads=threads, pooled=True), "reference"
    if not cpu_planners.have_oracle():
        cpu_planners.build("oracle")
    return cpu_planners.OraclePlanner(threads=threads, pooled=True), "port"


def time_cpu(sc, ticks, warmup=1, min_seconds=0.0):
    """Closed-loop ticks of the reference CPU path on all host threads; returns (steps/s, seconds, kind, cores, ticks run).
    min_seconds > 0: the `ticks`-long closed loop is repeated from a fresh plan (so that the agents never arrive and
    every tick keeps its A * (H - 1) steps) until that much CPU wall time has been sampled."""
    cores = host_threads()
    p, kind = cpu_planner(cores)
    steps = 0
    t_total = 0.0
    n_ticks = 0
    while True:
        feed = loop.ObstacleFeed(sc)
        loop.plan_begin(p, sc)
        for t in range(warmup + ticks):
            t0 = time.perf_counter()
            loop.control_tick(p, sc, feed)  # start_prediction runs the pooled rollout to termination
            p.stop_prediction()
            dt = time.perf_counter() - t0
            if t >= warmup:
                t_total += dt
                n_ticks += 1
                steps += int(p.get_agent_summaries()["steps"].sum()) - sc.num_agents
            feed.step()
        if t_total >= min_seconds:
            break
    p.close()
    return steps / t_total, t_total, kind, cores, n_ticks


def run_reference(args, rank, world):
    if rank != 0:
        return
    sc = WORKLOADS[args.workload]()
    ticks = max(1, args.steps)
    cores = host_threads()
    # bound the run: ~65 ns per (agent, step, obstacle) per core
    est = sc.num_agents * sc.max_prediction_steps * sc.num_obstacles * 65e-9 / max(cores, 1)
    budget = 120.0
    if est * (ticks + args.warmup) > budget:
        ticks = max(1, int(budget / est) - args.warmup)
    value, seconds, kind, cores, ticks = time_cpu(sc, ticks, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
            "steps": ticks, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / ticks, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(sc, 1),
            "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": kind,
                             "sample": f"{ticks} closed-loop control ticks of {sc.name}, all agents, pooled driver "
                                       f"over the reference's per-step methods on {cores} host threads"},
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- the CUDA arm -----------------------------------------------------------------------------------------------------
class Env:
    """Process-group plumbing of one rank (torch.distributed is used for nothing else)."""

    def __init__(self, rank, world, local_rank):
        import torch

        self.torch, self.rank, self.world, self.local_rank = torch, rank, world, local_rank
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxes, sums):
        t = self.torch.tensor(maxes, dtype=self.torch.float64, device="cuda")
        n = self.torch.tensor(sums, dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(n, op=self.dist.ReduceOp.SUM)
        return t.tolist(), n.tolist()

    def all_ok(self, ok):
        (worst,), _ = self.reduce([0.0 if ok else 1.0], [0.0])
        return worst == 0.0

    def make_manager(self, args):
        from pmaf_b200 import planner

        kw = dict(lanes_per_agent=args.lanes, block_threads=args.block, occupancy=args.occ)
        if self.world == 1:
            return planner.CfManager(self.local_rank, **kw)
        from pmaf_b200 import sharded

        return sharded.ShardedCfManager(self.local_rank, self.rank, self.world, **kw)


def measure(env, args, sc1, steps, warmup, seeded_random=False):
    """One workload: `warmup` untimed ticks, then `steps` device-timed ticks (value) and `steps` end-to-end
    ticks (e2e). Returns the record (identical on every rank; rank 0 prints it)."""
    world = env.world
    # weak scaling: every GPU owns sc1.num_agents agents of a population of sc1.num_agents * world
    sc = sc1 if world == 1 else sc1.with_(num_agents=sc1.num_agents * world,
                                          name=sc1.name.replace(f"_{sc1.num_agents}x", f"_{sc1.num_agents * world}x"))
    mgr = env.make_manager(args)
    feed = loop.ObstacleFeed(sc)
    if seeded_random:  # large sharded populations: the library draws each rank's vectors itself (shard-independent streams)
        mgr.seed_random_vecs(sc.seed)
    loop.plan_begin(mgr, sc, random_vecs=not seeded_random)
    env.barrier()  # ranks finish their (differently long) initialisation before the first exchange waits on a peer

    n_feed = sc.num_obstacles - 1 if feed.active else 0
    state = {"resident": False}

    def tick():
        # moving scenes: the obstacle feed (dynamic_obstacle_node's integration step) runs on the device-resident
        # list (pmaf_feed_obstacles, applied inside the tick kernel); the list crosses the bus once
        if state["resident"] and feed.active:
            out = mgr.tick(None, None, None, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist, sc.k_workspace,
                           sc.ws_limits)
        else:
            out = mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                           sc.k_workspace, sc.ws_limits)
            state["resident"] = True
        feed.step()  # the host copy stays in step: the e2e leg below passes it
        if feed.active:
            mgr.feed_obstacles(n_feed, feed.frequency)
        return out

    def timed_ticks(n):
        ms = 0.0
        for _ in range(n):
            mgr.flush_l2()
            mgr.stop_prediction()
            mgr.timer_start()
            tick()
            ms += mgr.timer_stop()  # waits for the rollout this tick launched
        return ms

    mgr.set_rollout_timing(False)  # production setting: no per-kernel events, the rollout is a dependent launch
    for _ in range(max(warmup, 3)):
        tick()
    mgr.stop_prediction()
    c0 = mgr.counters()
    env.barrier()
    dev_ms = timed_ticks(steps)
    env.barrier()
    c1 = mgr.counters()
    steps_local = c1["agent_steps_total"] - c0["agent_steps_total"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    # roofline pass: the same ticks again with CUDA events around every rollout kernel (they would serialise the
    # tick's two launches, so they stay out of the pass `value` comes from)
    mgr.set_rollout_timing(True)
    k0 = mgr.counters()
    timed_ticks(steps)
    mgr.stop_prediction()
    k1 = mgr.counters()
    rollout_ms = (k1["rollout_ms_total"] - k0["rollout_ms_total"]) / steps
    steps_per_launch = (k1["agent_steps_total"] - k0["agent_steps_total"]) / steps

    # ---- e2e: the reference-facing calls with host buffers, wall clock ----
    # the library's C++ host loop (pmaf_dry_run: planCallback's five CfManager calls per tick through the C ABI,
    # obstacle lists in host memory, uploaded by every call that takes them); L2 flushed before each tick and
    # the tick's rollout awaited inside its timed region
    mgr.set_upload_dedup(False)
    mgr.stop_prediction()
    e0 = mgr.counters()
    env.barrier()
    tick_s = []
    e2e_s, _, _, _ = mgr.dry_run(steps, feed.pos, feed.vel, feed.rad, n_feed, sc.delta_t, sc.k_goal_dist,
                                 sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits,
                                 feed_frequency=feed.frequency, wait_rollout=True, flush_l2=True, tick_times=tick_s)
    e1 = mgr.counters()
    e2e_steps_local = e1["agent_steps_total"] - e0["agent_steps_total"]
    (dev_ms, e2e_s, rollout_ms), (steps_all, e2e_steps_all) = env.reduce([dev_ms, e2e_s, rollout_ms],
                                                                         [steps_local, e2e_steps_local])
    O = sc.num_obstacles
    alg_bytes = 24.0 * steps_per_launch + 128.0 * mgr.A + 56.0 * O
    alg_flops = (40.0 * (O - 1) + 100.0) * steps_per_launch
    rec = {
        "value": steps_all / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps, "steps": steps, "warmup": max(warmup, 3),
        "config": workload_config(sc, world, {"lanes_per_agent": c1["lanes_per_agent"],
                                              "block_threads": c1["block_threads"], "grid": c1["grid_blocks"],
                                              "smem_bytes": c1["smem_bytes"], "occupancy_build": c1["occupancy_build"],
                                              "best_agent_exchange": getattr(mgr, "exchange", "none (one GPU)")}),
        "e2e": {"value": e2e_steps_all / e2e_s, "unit": "agent-steps/s",
                "h2d_bytes_per_step": (e1["h2d_bytes"] - e0["h2d_bytes"]) / steps,
                "d2h_bytes_per_step": (e1["d2h_bytes"] - e0["d2h_bytes"]) / steps,
                "ms_per_step": 1e3 * e2e_s / steps,
                "ms_per_step_median_max_rank0": [1e3 * float(np.median(tick_s)), 1e3 * float(np.max(tick_s))],
                "api": "pmaf_dry_run (C++ host loop): per tick stop_prediction, evaluate_agents, move_real_agent, "
                       "get_next_position/velocity, reset_agents, start_prediction on host obstacle lists"},
        "gpu_launches": int(launches),
        "_kernel": {"ms": rollout_ms, "alg_bytes": alg_bytes, "alg_flops": alg_flops,
                    "general_step_share": (c1["general_steps_total"] - c0["general_steps_total"]) / max(steps_local, 1),
                    "timing": "CUDA events around the rollout kernel in a second pass over the same ticks"},
    }
    return rec, mgr


def roofline(rec, fp64_peak, peaks, peak_src, traffic):
    """The roofline that binds is the FP64 pipe (~850 flop per byte at O = 256 against a machine balance
    of ~5 flop/B for binary64), so `frac` is algorithmic binary64 flops / kernel time / the measured DFMA peak; the HBM
    figure the contract asks for rides along under `hbm`."""
    k = rec.pop("_kernel")
    tf = k["alg_flops"] / (k["ms"] * 1e-3) / 1e12
    gbs = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9
    return {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
            "traffic": traffic, "peak_source": "measured in this run: independent DFMA chains on all SMs (pmaf_measure_fp64_peak)",
            "algorithmic_flops": k["alg_flops"], "algorithmic_flops_per_agent_step": "40*(O-1)+100 (SURVEY.md §8d)",
            "kernel": "rollout_kernel", "kernel_ms": k["ms"], "kernel_ms_timing": k["timing"],
            "kernel_share_of_step": k["ms"] / rec["ms_per_step"],
            "general_step_share": k["general_step_share"],
            "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                    "algorithmic_bytes": k["alg_bytes"], "peak_source": peak_src,
                    "note": "24 B per agent-step + 128 B per agent + 56 B per obstacle: the path is not HBM-bound"}}


def sharded_parity(env, args, name="near326_switching", ticks=3):
    """N > 1: `ticks` control ticks of a golden case sharded over the ranks; best ids and the real agent's
    states (replicated) and this rank's block of per-agent results must equal the reference's golden vectors
    bit for bit. Returns "ok" or a description of the first mismatch (collective: every rank calls it)."""
    from pmaf_b200 import cases, sharded

    want = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sc = cases.all_cases()[name].scenario
    mgr = env.make_manager(args)
    got = loop.run_closed_loop(mgr, sc, ticks)
    first, end = sharded.shard_range(sc.num_agents, env.rank, env.world)
    bad = []

    def same(a, b):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        if a.shape != b.shape:
            return False
        if b.dtype.kind == "f":
            return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))
        return bool(np.all(a == b))

    for k in ("best", "next_pos", "next_vel", "goal_dist"):
        if not same(got[k], want[k][:ticks]):
            bad.append(k)
    for k in ("steps", "length", "min_obs_dist", "reached"):
        if not same(got[k], want[k][:ticks, first:end]):
            bad.append(k)
    if mgr.counters()["collectives"] < ticks:
        bad.append("no exchange ran")
    mgr.close()
    ok = env.all_ok(not bad)
    return "ok" if ok else f"MISMATCH (rank {env.rank}: {bad or 'another rank'})"


def run_ours(args, rank, world, local_rank):
    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner comes out on
    # fd 1 whatever NCCL_DEBUG_FILE says) goes to stderr; the line itself is written to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libpmaf has no CPU fallback")
    torch.cuda.set_device(local_rank)
    env = Env(rank, world, local_rank)
    parity = sharded_parity(env, args) if world > 1 else None
    if parity not in (None, "ok"):
        raise SystemExit(f"bench.py: sharded parity failed before timing: {parity}")
    sampler = ClockSampler(local_rank).wait_ready() if rank == 0 else None
    rec, mgr = measure(env, args, WORKLOADS[args.workload](), args.steps, args.warmup)
    fp64_peak = mgr.measure_fp64_peak()
    mgr.close()
    peaks, peak_src = measured_peaks()
    line = {"metric": METRIC, "value": rec["value"], "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": rec["config"], "e2e": rec["e2e"],
            "gpu_launches": rec["gpu_launches"],
            "roofline": roofline(rec, fp64_peak, peaks, peak_src, measured_traffic(args.workload) if world == 1 else None)}
    if parity:
        line["sharded_parity"] = parity
    # ---- sub-records: the other BASELINE configs, time-boxed ----
    subs = []
    if not args.no_sub and args.workload == "c2":
        names = ["c3", "c5", "c4s"] if world == 1 else ["c4s"]
        for name in names:
            label = name if world == 1 else "c4"
            try:  # a failing sub-record must not take the headline line with it
                r, m = measure(env, args, WORKLOADS[name](), SUB_TICKS, 3, seeded_random=(world > 1))
                m.close()
                r["roofline"] = roofline(r, fp64_peak, peaks, peak_src, measured_traffic(name) if world == 1 else None)
                r.update({"name": label, "metric": METRIC, "unit": "agent-steps/s", "n_gpus": world})
            except Exception as e:  # noqa: BLE001
                r = {"name": label, "error": f"{type(e).__name__}: {e}"}
            subs.append(r)
    line["workloads"] = subs
    line["clocks"] = sampler.stop() if sampler else None
    if rank == 0:
        if world == 1 and not args.no_cpu:
            sc = WORKLOADS[args.workload]()
            cores = host_threads()
            est = sc.num_agents * sc.max_prediction_steps * sc.num_obstacles * 65e-9 / max(cores, 1)
            ticks = int(min(50, max(2, 15.0 / max(est, 1e-6))))
            v, seconds, kind, cores, n_ticks = time_cpu(sc, ticks, min_seconds=10.0)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": kind,
                                    "sample": f"{n_ticks} closed-loop control ticks of {sc.name} (all "
                                              f"{sc.num_agents} agents; the {ticks}-tick loop repeated from a fresh "
                                              f"plan), pooled driver on {cores} host threads, {seconds:.1f} s of CPU "
                                              f"wall time"}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if env.dist:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--occ", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records of the other BASELINE configs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
