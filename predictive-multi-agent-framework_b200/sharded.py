"""Agent sharding across the GPUs of one node.

Agents are independent within a control tick (cf_manager.cpp:118-123: one thread each), so rank r
owns the contiguous block [r*A/W, (r+1)*A/W) of the global population; agent types and gains follow
the GLOBAL index. Obstacles and the real agent are replicated (every rank steps its own replica of
the real agent — deterministic, so the replicas stay bit-identical). The only exchange is the
best-agent selection: per evaluate ONE record per rank

    { min_cost, incumbent_cost, cost_agent0, min_index, owns_incumbent, random_vecs[O][3] }

(40 + 24*O bytes), stored by every rank straight into every peer's cudaIpc-mapped exchange block over
NVLink inside the tick kernel (local scan, exchange, replicated selection and the real agent's step are
one launch), or — when peer mapping is unavailable — gathered by one NCCL all-gather that libpmaf
enqueues on the planner's stream; no host round trip either way. `local_record` /
`select_global_best` below restate that protocol on the host; the gloo tests use them to check
that the sharded selection equals the reference's serial scan over the whole population.

torch.distributed is plumbing only: it distributes the NCCL unique id and the cudaIpc handles (and, in
bench.py, the timing reductions).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .planner import CfManager

DBL_MAX = float(np.finfo(np.float64).max)
NO_INDEX = 0x7FFFFFFF


def shard_range(n_global, rank, world):
    """Global agent indices [first, end) owned by `rank`."""
    return rank * n_global // world, (rank + 1) * n_global // world


def local_record(costs_local, first_agent, incumbent_id):
    """Host restatement of evaluate_kernel's per-rank record (csrc/pmaf_rollout.cuh): serial scan with
    strict '<' starting from DBL_MAX (cf_manager.cpp:335-342); incumbent_id is the reference's agent id
    (global index + 1) or 0 when there is no incumbent."""
    min_cost, min_index = DBL_MAX, NO_INDEX
    for i, c in enumerate(costs_local):
        if c < min_cost:
            min_cost, min_index = float(c), first_agent + i
    inc = incumbent_id - 1 - first_agent if incumbent_id > 0 else -1
    owns = 0 <= inc < len(costs_local)
    return dict(min_cost=min_cost, min_index=min_index, owns_incumbent=owns,
                incumbent_cost=float(costs_local[inc]) if owns else math.nan,
                cost_agent0=float(costs_local[0]) if first_agent == 0 and len(costs_local) else math.nan)


def select_global_best(records, incumbent_id):
    """Host restatement of global_select_kernel: scan the ranks in order (contiguous ascending blocks
    keep 'lowest index wins'), then the 0.9 hysteresis (cf_manager.cpp:343-350).
    Returns (best_index, new_incumbent_id)."""
    min_cost, min_idx, min_rank = DBL_MAX, 0, -1
    inc_cost, have_inc = 0.0, False
    for r, rec in enumerate(records):
        if rec["min_index"] != NO_INDEX and rec["min_cost"] < min_cost:
            min_cost, min_idx, min_rank = rec["min_cost"], rec["min_index"], r
        if rec["owns_incumbent"]:
            inc_cost, have_inc = rec["incumbent_cost"], True
    if min_rank < 0:
        min_cost = records[0]["cost_agent0"]
    if incumbent_id > 0 and have_inc and not (min_cost < 0.9 * inc_cost):
        return incumbent_id - 1, incumbent_id
    return min_idx, min_idx + 1


def serial_reference_selection(costs, incumbent_id):
    """CfManager::evaluateAgents' selection over the whole population (cf_manager.cpp:334-355)."""
    min_idx, min_cost = 0, DBL_MAX
    for i, c in enumerate(costs):
        if c < min_cost:
            min_cost, min_idx = c, i
    if incumbent_id > 0 and incumbent_id - 1 < len(costs):
        if costs[min_idx] < 0.9 * costs[incumbent_id - 1]:
            return min_idx, min_idx + 1
        return incumbent_id - 1, incumbent_id
    return min_idx, min_idx + 1


def exchange_nccl_id(make_id, rank, src=0, group=None):
    """Rank `src` creates the 128-byte NCCL unique id; everyone receives it over torch.distributed
    (any backend: gloo on CPU, nccl on GPU)."""
    import torch.distributed as dist

    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    ident = bytes(box[0])
    assert len(ident) == 128
    return ident


class ShardedCfManager(CfManager):
    """CfManager whose agents are one rank's block of a population sharded over `world` GPUs.
    Same call surface; per-agent getters return the LOCAL block, evaluate/tick return GLOBAL indices."""

    def __init__(self, device, rank, world, group=None, p2p=True, **tuning):
        super().__init__(device, **tuning)
        self.rank, self.world = int(rank), int(world)

        def make_id():
            buf = C.create_string_buffer(128)
            self._check(self.lib.pmaf_nccl_unique_id(buf))
            return buf.raw

        ident = exchange_nccl_id(make_id, self.rank, group=group)
        self._check(self.lib.pmaf_nccl_init(self.h, ident, self.rank, self.world))
        self.exchange = "nccl"
        if p2p and self.world <= 16:
            self._try_p2p(group)

    def _try_p2p(self, group):
        """Peer-memory exchange (pmaf_p2p_export / pmaf_p2p_import): every rank's cudaIpc handle goes to every
        rank over torch.distributed; all ranks switch to it or all stay on NCCL."""
        import torch.distributed as dist

        buf = C.create_string_buffer(64)
        ok = self.lib.pmaf_p2p_export(self.h, buf) == 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, buf.raw), group=group)
        if ok and all(g[0] for g in gathered):
            handles = b"".join(g[1] for g in gathered)
            ok = self.lib.pmaf_p2p_import(self.h, handles, self.rank, self.world) == 0
        else:
            ok = False
        agreed = [None] * self.world
        dist.all_gather_object(agreed, ok, group=group)
        if all(agreed):
            self.exchange = "p2p"
        elif ok:  # someone could not map a peer: everybody back to NCCL
            self._check(self.lib.pmaf_p2p_export(self.h, buf))  # re-export drops the imported mappings

    def init(self, goal, delta_t, obs_pos, obs_vel, obs_rad, k_attr, *args, **kw):
        n_global = max(len(np.atleast_1d(k_attr)), 1)
        first, end = shard_range(n_global, self.rank, self.world)
        if end <= first:
            raise ValueError(f"rank {self.rank} of {self.world} owns no agent of {n_global}")
        self.set_shard(n_global, first, self.rank, self.world)
        super().init(goal, delta_t, obs_pos, obs_vel, obs_rad, k_attr, *args, **kw)
