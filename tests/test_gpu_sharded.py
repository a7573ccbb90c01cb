"""Sharded planner on 2+ GPUs (gpurun --gpus 2): every rank owns a contiguous block of agents, one
best-agent exchange per tick (P2P stores into the peers' exchange blocks fused with the selection, or
the NCCL all-gather fallback) selects the global best. The sharded run must be bit-identical to the
reference goldens of the unsharded population: best ids, real-agent trajectory, and each rank's
block of paths — through the call-by-call API and through the fused tick (pmaf_tick: exchange and
selection inside tick_kernel, the rollout as its programmatic dependent)."""
import os

import numpy as np
import pytest

from mp_util import run_ranks

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _init(rank, world, init_file):
    """Device + process group of one rank (gloo over a file store: plumbing only, see mp_util)."""
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pmaf_b200  # noqa: F401
    import torch
    from mp_util import init_gloo

    torch.cuda.set_device(rank)
    return init_gloo(rank, world, init_file)


def _worker(rank, world, init_file, q, name, p2p=True):
    dist = _init(rank, world, init_file)
    from pmaf_b200 import cases, sharded

    try:
        case = cases.all_cases()[name]
        mgr = sharded.ShardedCfManager(rank, rank, world, p2p=p2p)
        got = case(mgr)
        want = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        first, end = sharded.shard_range(case.scenario.num_agents, rank, world)

        def same(a, b):
            a, b = np.asarray(a), np.asarray(b)
            return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b)) if b.dtype.kind == "f" else a == b))

        errs = []
        for k in ("best", "next_pos", "next_vel", "trajectory", "best_type", "best_id", "goal_dist"):
            if k in want and not same(got[k], want[k]):
                errs.append(k)
        for k in ("steps", "length", "min_obs_dist", "reached"):  # [ticks, A] -> this rank's columns
            if not same(got[k], want[k][..., first:end]):
                errs.append(k)
        if not same(got["final_paths"], want["final_paths"][first:end]):
            errs.append("final_paths")
        if not same(got["known"][:-1], want["known"][first:end]) or not same(got["known"][-1], want["known"][-1]):
            errs.append("known")
        c = mgr.counters()
        if c["collectives"] < len(want["best"]):
            errs.append(f"collectives={c['collectives']}")
        if mgr.exchange != ("p2p" if p2p else "nccl"):
            errs.append(f"exchange={mgr.exchange}")
        q.put((rank, errs))
        mgr.close()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, [repr(e), traceback.format_exc()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,p2p", [("near326_switching", True), ("near326_reinit_random_incumbent", True),
                                      ("rand5_many_agents", True), ("moving1", True), ("near326_switching", False)])
def test_sharded_equals_unsharded_reference(name, p2p):
    """p2p: best-agent exchange by P2P stores into the peers' cudaIpc-mapped blocks (one fused kernel);
    otherwise the NCCL all-gather."""
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    results = run_ranks(_worker, world, args=(name, p2p), timeout=240)
    assert len(results) == world and all(r is not None and not errs for r, errs in results), results


def _worker_tick(rank, world, init_file, q, name, p2p, timing):
    dist = _init(rank, world, init_file)
    from pmaf_b200 import cases, loop, sharded

    try:
        sc = cases.all_cases()[name].scenario
        want = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        ticks = min(len(want["best"]), 40)
        mgr = sharded.ShardedCfManager(rank, rank, world, p2p=p2p)
        mgr.set_rollout_timing(timing)
        feed = loop.ObstacleFeed(sc)
        loop.plan_begin(mgr, sc)
        first, end = sharded.shard_range(sc.num_agents, rank, world)
        errs = []
        for t in range(ticks):
            b, x, v = mgr.tick(feed.pos, feed.vel, feed.rad, sc.delta_t, sc.k_goal_dist, sc.k_path_len, sc.k_safe_dist,
                               sc.k_workspace, sc.ws_limits)
            mgr.stop_prediction()
            s = mgr.get_agent_summaries()
            if b != want["best"][t] or not np.array_equal(x, want["next_pos"][t]) or not np.array_equal(v, want["next_vel"][t]):
                errs.append(f"tick {t}: best {b} vs {want['best'][t]}")
                break
            if not np.array_equal(s["steps"], want["steps"][t, first:end]) or not np.array_equal(s["length"], want["length"][t, first:end]):
                errs.append(f"tick {t}: agent block differs")
                break
            feed.step()
        q.put((rank, errs))
        mgr.close()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, [repr(e), traceback.format_exc()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,p2p,timing", [("near326_switching", True, False), ("near326_switching", True, True),
                                             ("moving1", True, False), ("near326_switching", False, False)])
def test_sharded_fused_tick_equals_reference(name, p2p, timing):
    """pmaf_tick on a sharded planner: local scan, exchange and replicated selection inside tick_kernel."""
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    results = run_ranks(_worker_tick, world, args=(name, p2p, timing), timeout=240)
    assert len(results) == world and all(r is not None and not errs for r, errs in results), results


def _worker_c4(rank, world, init_file, q, ticks, agents_per_rank, shm_dir):
    """BASELINE configs[3]: 65 536 agents x 1024 obstacles x horizon 200 sharded over 8 ranks (or the same shape at
    agents_per_rank * world agents), every tick of every agent against the C oracle run over the WHOLE population on
    rank 0's host cores: best ids, real-agent states and, per agent, executed steps, path length (an order-sensitive
    sum over every path point), minimum obstacle distance and goal flag — bit for bit."""
    dist = _init(rank, world, init_file)
    from pmaf_b200 import loop, scenarios, sharded

    try:
        sc = scenarios.c4(agents_per_rank * world)
        ref_file = os.path.join(shm_dir, "c4_oracle.npz")
        if rank == 0:
            from oracle import cpu_planners

            if not cpu_planners.have_oracle():
                cpu_planners.build("oracle")
            orc = cpu_planners.OraclePlanner(threads=len(os.sched_getaffinity(0)), pooled=True)
            want = loop.run_closed_loop(orc, sc, ticks)
            orc.close()
            np.savez(ref_file, **want)
        mgr = sharded.ShardedCfManager(rank, rank, world)
        got = loop.run_closed_loop(mgr, sc, ticks)
        dist.barrier()
        want = np.load(ref_file)
        first, end = sharded.shard_range(sc.num_agents, rank, world)
        errs = []

        def same(a, b):
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
            if a.shape != b.shape:
                return False
            if b.dtype.kind == "f":
                return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))
            return bool(np.all(a == b))

        for k in ("best", "next_pos", "next_vel", "goal_dist"):
            if not same(got[k], want[k]):
                errs.append(k)
        for k in ("steps", "length", "min_obs_dist", "reached"):
            if not same(got[k], want[k][:, first:end]):
                errs.append(k)
        if int(got["steps"][-1].sum()) != (end - first) * sc.max_prediction_steps:
            errs.append("not every step executed")
        if mgr.exchange != "p2p":
            errs.append(f"exchange={mgr.exchange}")
        q.put((rank, errs))
        mgr.close()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, [repr(e), traceback.format_exc()]))
    finally:
        dist.destroy_process_group()


def test_c4_full_population_over_8_ranks_matches_oracle():
    import tempfile

    import torch

    world = torch.cuda.device_count()
    if world < 8:
        pytest.skip("needs 8 GPUs (gpurun --gpus 8): BASELINE configs[3], 65 536 agents over 8 ranks")
    world = 8
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as shm_dir:
        results = run_ranks(_worker_c4, world, args=(2, 8192, shm_dir), timeout=900)
    assert len(results) == world and all(r is not None and not errs for r, errs in results), results
