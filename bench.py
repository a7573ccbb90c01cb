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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def cpu_planner(threads=0):
    from oracle import cpu_planners

    if cpu_planners.have_ref():
        return cpu_planners.RefPlanner(threads=threads, pooled=True), "reference"
    if not cpu_planners.have_oracle():
        cpu_planners.build("oracle")
    return cpu_planners.OraclePlanner(threads=threads, pooled=True), "port"


def time_cpu(sc, ticks, warmup=1):
    """Closed-loop ticks of the reference CPU path on all host threads; returns (steps/s, seconds, kind, cores)."""
    p, kind = cpu_planner()
    cores = p.host_threads()
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(p, sc)
    steps = 0
    t_total = 0.0
    for t in range(warmup + ticks):
        t0 = time.perf_counter()
        loop.control_tick(p, sc, feed)  # start_prediction runs the pooled rollout to termination
        p.stop_prediction()
        dt = time.perf_counter() - t0
        if t >= warmup:
            t_total += dt
            steps += int(p.get_agent_summaries()["steps"].sum()) - sc.num_agents
        feed.step()
    p.close()
    return steps / t_total, t_total, kind, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    sc = WORKLOADS[args.workload]()
    ticks = max(1, args.steps)
    # bound the run: ~65 ns per (agent, step, obstacle) per core
    est = sc.num_agents * sc.max_prediction_steps * sc.num_obstacles * 65e-9 / max(os.cpu_count() or 1, 1)
    budget = 120.0
    if est * (ticks + args.warmup) > budget:
        ticks = max(1, int(budget / est) - args.warmup)
    value, seconds, kind, cores = time_cpu(sc, ticks, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
            "steps": ticks, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / ticks, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(sc, 1),
            "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": kind,
                             "sample": f"{ticks} closed-loop control ticks of {sc.name}, all agents, pooled driver "
                                       f"over the reference's per-step methods on {cores} host threads"},
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pmaf_b200 import planner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libpmaf has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sc1 = WORKLOADS[args.workload]()
    # weak scaling: every GPU owns `agents` agents of a population of agents * world
    sc = sc1 if world == 1 else sc1.with_(num_agents=sc1.num_agents * world, name=f"{sc1.name}_x{world}")
    if world == 1:
        mgr = planner.CfManager(local_rank, lanes_per_agent=args.lanes, block_threads=args.block, occupancy=args.occ)
    else:
        from pmaf_b200 import sharded

        mgr = sharded.ShardedCfManager(local_rank, rank, world, lanes_per_agent=args.lanes, block_threads=args.block,
                                       occupancy=args.occ)
    feed = loop.ObstacleFeed(sc)
    loop.plan_begin(mgr, sc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tick():
        out = mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                       sc.k_workspace, sc.ws_limits)
        feed.step()
        return out

    for _ in range(max(args.warmup, 3)):
        tick()
    mgr.stop_prediction()
    c0 = mgr.counters()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms = 0.0
    for _ in range(args.steps):
        mgr.flush_l2()
        mgr.stop_prediction()
        mgr.timer_start()
        tick()
        dev_ms += mgr.timer_stop()  # waits for the rollout this tick launched
    barrier()
    c1 = mgr.counters()
    steps_local = c1["agent_steps_total"] - c0["agent_steps_total"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    rollout_ms = (c1["rollout_ms_total"] - c0["rollout_ms_total"]) / args.steps

    # ---- e2e: the reference-facing calls with host buffers, wall clock ----
    # the library's C++ host loop (pmaf_dry_run: planCallback's five CfManager calls per tick through the C ABI,
    # obstacle lists in host memory, uploaded by every call that takes them); L2 flushed before each tick and
    # the tick's rollout awaited inside its timed region
    mgr.set_upload_dedup(False)
    mgr.stop_prediction()
    e0 = mgr.counters()
    barrier()
    n_feed = sc.num_obstacles - 1 if feed.active else 0
    e2e_s, _, _, _ = mgr.dry_run(args.steps, feed.pos, feed.vel, feed.rad, n_feed, sc.delta_t, sc.k_goal_dist,
                                 sc.k_path_len, sc.k_safe_dist, sc.k_workspace, sc.ws_limits,
                                 feed_frequency=feed.frequency, wait_rollout=True, flush_l2=True)
    e1 = mgr.counters()
    e2e_steps_local = e1["agent_steps_total"] - e0["agent_steps_total"]
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions (device-resident and e2e)

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    n = torch.tensor([steps_local, e2e_steps_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = t.tolist()
    steps_all, e2e_steps_all = n.tolist()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = steps_all / (dev_ms * 1e-3)
        O = sc.num_obstacles
        steps_per_launch = steps_local / args.steps
        alg_bytes = 24.0 * steps_per_launch + 128.0 * mgr.A + 56.0 * O
        alg_flops = (40.0 * (O - 1) + 100.0) * steps_per_launch
        hbm_achieved = alg_bytes / (rollout_ms * 1e-3) / 1e9
        fp64_peak = mgr.measure_fp64_peak()
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(sc, world, {"lanes_per_agent": c1["lanes_per_agent"],
                                                  "block_threads": c1["block_threads"], "grid": c1["grid_blocks"],
                                                  "smem_bytes": c1["smem_bytes"], "occupancy_build": c1["occupancy_build"],
                                                  "best_agent_exchange": getattr(mgr, "exchange", "none (one GPU)")}),
            "clocks": clocks,
            "e2e": {"value": e2e_steps_all / e2e_s, "unit": "agent-steps/s",
                    "h2d_bytes_per_step": (e1["h2d_bytes"] - e0["h2d_bytes"]) / args.steps,
                    "d2h_bytes_per_step": (e1["d2h_bytes"] - e0["d2h_bytes"]) / args.steps,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "pmaf_dry_run (C++ host loop): per tick stop_prediction, evaluate_agents, move_real_agent, "
                           "get_next_position/velocity, reset_agents, start_prediction on host obstacle lists"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": hbm_achieved / peaks["hbm_gbs"], "traffic": measured_traffic(args.workload) if world == 1 else None,
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "kernel": "rollout_kernel", "kernel_ms": rollout_ms,
                         "kernel_share_of_step": rollout_ms / (dev_ms / args.steps),
                         "note": "the path is FP64-issue/latency bound, not HBM bound (SURVEY.md §8d); "
                                 "see fp64_pipe for the roofline that binds",
                         "fp64_pipe": {"achieved": alg_flops / (rollout_ms * 1e-3) / 1e12, "peak": fp64_peak,
                                       "unit": "TFLOP/s",
                                       "frac": alg_flops / (rollout_ms * 1e-3) / 1e12 / fp64_peak,
                                       "peak_source": "measured here: dependent DFMA chains on all SMs"}},
        }
        if world == 1 and not args.no_cpu:
            est = sc.num_agents * sc.max_prediction_steps * O * 65e-9 / max(os.cpu_count() or 1, 1)
            ticks = int(min(50, max(2, 15.0 / max(est, 1e-6))))
            v, seconds, kind, cores = time_cpu(sc, ticks)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": kind,
                                    "sample": f"{ticks} closed-loop control ticks of {sc.name} (all "
                                              f"{sc.num_agents} agents), pooled driver on {cores} host threads, "
                                              f"{seconds:.1f} s"}
        print(json.dumps(line), flush=True)
    mgr.close()
    if world > 1:
        dist.destroy_process_group()


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
