"""Helpers for the multi-process tests (gloo on CPU, sharded planners on GPUs).

Rendezvous goes through a FILE store (`init_method=file://...`), not a TCP port: picking a "free" port and
handing it to the ranks is a race (another process can take it in between — seen as EADDRINUSE on a busy box),
and a rank that dies in rendezvous leaves the others waiting. `run_ranks` also makes sure no rank outlives a
failed test: the children are daemons and are terminated before the results are judged."""
import os
import queue as queue_mod
import tempfile


def init_gloo(rank, world, init_file):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    return dist


def run_ranks(target, world, args=(), timeout=300):
    """Spawn `world` processes running target(rank, world, init_file, result_queue, *args); returns the list of
    (rank, payload) every rank put on the queue. A rank that does not report in time is reported as
    (None, "timeout") and all ranks are terminated."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    results = []
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        procs = [ctx.Process(target=target, args=(r, world, init_file, q, *args), daemon=True) for r in range(world)]
        for p in procs:
            p.start()
        try:
            for _ in procs:
                try:
                    results.append(q.get(timeout=timeout))
                except queue_mod.Empty:
                    results.append((None, "timeout: a rank never reported"))
                    break
        finally:
            for p in procs:
                p.join(timeout=20 if len(results) == world and all(r[0] is not None for r in results) else 1)
            for p in procs:
                if p.is_alive():
                    p.terminate()
            for p in procs:
                p.join(timeout=10)
    return results
