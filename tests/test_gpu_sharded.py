"""Sharded planner on 2+ GPUs (gpurun --gpus 2): every rank owns a contiguous block of agents, one
best-agent exchange per tick (P2P stores into the peers' exchange blocks fused with the selection, or
the NCCL all-gather fallback) selects the global best. The sharded run must be bit-identical to the
reference goldens of the unsharded population: best ids, real-agent trajectory, and each rank's
block of paths — through the call-by-call API and through the fused tick (pmaf_tick: exchange and
selection inside tick_kernel, the rollout as its programmatic dependent)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, q, p2p=True):
    import sys

    sys.path.insert(0, ROOT)
    import pmaf_b200  # noqa: F401
    import torch
    import torch.distributed as dist
    from pmaf_b200 import cases, sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
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
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(not errs for _, errs in results), results


def _worker_tick(rank, world, port, name, q, p2p, timing):
    import sys

    sys.path.insert(0, ROOT)
    import pmaf_b200  # noqa: F401
    import torch
    import torch.distributed as dist
    from pmaf_b200 import cases, loop, sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
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
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_tick, args=(r, world, port, name, q, p2p, timing)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(not errs for _, errs in results), results
