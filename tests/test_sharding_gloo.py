"""world_size-2 (and 3) gloo tests, on CPU, of the host-side logic of the sharded planner: block
partitioning, the NCCL-id exchange plumbing, and the best-agent exchange protocol (per-rank record
-> all-gather -> replicated serial selection), which must equal the reference's serial argmin with
hysteresis over the whole population — ties, NaN costs, incumbents on other ranks included."""
import numpy as np
import pytest

from mp_util import init_gloo, run_ranks
from pmaf_b200 import sharded


def _worker(rank, world, init_file, q):
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist = init_gloo(rank, world, init_file)
    try:
        # 1. the id exchange: every rank ends up with rank 0's 128 bytes
        ident = sharded.exchange_nccl_id(lambda: bytes(range(128)), rank)
        assert ident == bytes(range(128))
        # 2. selection protocol on seeded synthetic costs, many trials
        rng = np.random.default_rng(1234)  # same stream on every rank
        for trial in range(300):
            n = int(rng.integers(world, 40))
            costs = rng.choice([1.0, 2.0, 3.5, 100.0], n) + rng.choice([0.0, 0.0, 1e-9], n)
            if trial % 7 == 0:
                costs[rng.integers(0, n)] = np.nan
            if trial % 11 == 0:
                costs[:] = np.nan
            if trial % 13 == 0:
                costs[:] = 5.0  # all equal: index 0 wins (first tick, quirk 7)
            incumbent_id = int(rng.integers(0, n + 1))  # 0 = none
            first, end = sharded.shard_range(n, rank, world)
            rec = sharded.local_record(costs[first:end], first, incumbent_id)
            gathered = [None] * world
            dist.all_gather_object(gathered, rec)
            got = sharded.select_global_best(gathered, incumbent_id)
            want = sharded.serial_reference_selection(costs, incumbent_id)
            assert got == want, (trial, rank, costs.tolist(), incumbent_id, got, want)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_selection_equals_serial_scan(world):
    results = run_ranks(_worker, world, timeout=120)
    assert len(results) == world and all(msg == "ok" for _, msg in results), results


def test_shard_ranges_partition_the_population():
    for n in (1, 5, 7, 256, 65536, 1000003):
        for w in (1, 2, 3, 4, 8):
            if n < w:
                continue
            r = [sharded.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - f for f, e in r]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
