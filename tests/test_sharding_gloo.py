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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_protocol_model_never_reads_an_overwritten_slot(world):
    """Model of the peer-memory exchange (csrc/pmaf_rollout.cuh: p2p_select_body): every rank stores its record into
    slot [parity][rank] of every peer's block, publishes the tick's sequence number per peer, waits for all flags of
    the tick in its own block, then reads all slots. Two buffers alternate by tick parity. The claim behind it: a rank
    can be at most one tick ahead of any other (its next selection needs everyone's next record), so a slot is never
    overwritten before its reader is done. Ranks are threads with random skew; every read must carry the reader's
    own sequence number."""
    import random
    import threading
    import time

    ticks = 300
    slots = [[[(-1, -1)] * world for _ in range(2)] for _ in range(world)]  # [owner block][parity][writer rank] = (seq, payload)
    flags = [[[0] * world for _ in range(2)] for _ in range(world)]
    errors = []

    def rank_main(r):
        rng = random.Random(1000 + r)
        for t in range(ticks):
            seq = t + 1
            parity = seq & 1
            if rng.random() < 0.2:
                time.sleep(rng.random() * 1e-3)  # skew: some ranks arrive late, others race ahead
            for peer in range(world):  # 1. the record into every rank's block
                slots[peer][parity][r] = (seq, 1000 * r + t)
            for peer in range(world):  # 2. publish
                flags[peer][parity][r] = seq
            t0 = time.time()
            while any(flags[r][parity][w] != seq for w in range(world)):  # 3. wait for everyone's record of this tick
                if time.time() - t0 > 20:
                    errors.append((r, t, "timeout"))
                    return
                time.sleep(0)
            for w in range(world):  # 4. the replicated selection reads every slot
                got = slots[r][parity][w]
                if got != (seq, 1000 * w + t):
                    errors.append((r, t, w, got))
                    return

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errors, errors[:5]
