"""Query-sharded overlap across the GPUs of one box.

One process per GPU (`torch.distributed`, backend nccl; gloo in the CPU tests).  The
query set is split into contiguous row ranges, one per rank; the smaller target
set is replicated with ONE broadcast from `src` (NCCL over NVLink); every rank
runs the single-GPU engine on (its queries) x (all targets); the only other
communication is an all-gather of the per-rank row counts so that every rank
knows the offset of its slice in the concatenated result.  Any partition of the
queries is valid because every query-side quantity depends only on the
replicated targets.

Row order of the concatenated result: rank-major, inside a rank the reference's
order for that shard.  For how='left' with keep_order=True (the default of
`bioframe.overlap`) `overlap_ordered_sharded` yields every shard in (query row,
target row) order, so the concatenation is the single-GPU result bit for bit; the
same holds for the other per-query results (coverage, count_overlaps, subtract,
closest per query).  For the block order of how='inner' the pair SET is identical
and the order is (shard, group, A-block, B-block)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, balanced row ranges: rank r owns [bounds[r], bounds[r+1])."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def broadcast_targets(tensors, src=0, group=None):
    """Replicate the target columns (same shapes allocated on every rank) from `src`."""
    for t in tensors:
        if t is not None:
            dist.broadcast(t, src=src, group=group)
    return tensors


def overlap_sharded(grp1, start1, end1, row_offset, grp2, start2, end2, n_groups, how="left", engine=None,
                    group=None, broadcast=True, src=0):
    """This rank's slice of `_overlap_intidxs` over the union of all ranks' queries.

    grp1/start1/end1: this rank's query shard (rows [row_offset, row_offset + n_local) of the full set).
    grp2/start2/end2: target columns; valid on `src`, allocated (any content) elsewhere when broadcast=True.
    Returns (events1 with GLOBAL query row numbers, events2, offset of this slice in the concatenated
    result, total rows over all ranks)."""
    if engine is None:
        from . import engine as engine  # the CUDA engine; tests inject an oracle-backed double
    if how in ("right", "outer"):
        raise ValueError("query-sharded overlap supports how='inner' and how='left' (unpaired targets need a reduction)")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world > 1 and broadcast:
        broadcast_targets([grp2, start2, end2], src=src, group=group)
    # the engine adds the shard's position to the query row numbers while it writes them
    ev1, ev2, _ = engine.overlap_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, how=how,
                                         row_offset1=int(row_offset))
    ev1 = torch.as_tensor(ev1)
    ev2 = torch.as_tensor(ev2)
    n_local = torch.tensor([ev1.numel()], dtype=torch.int64, device=ev1.device)
    if world > 1:
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        counts = torch.stack(counts).flatten().tolist()  # one device->host read
    else:
        counts = [int(n_local.item())]
    return ev1, ev2, int(sum(counts[:rank])), int(sum(counts))


def _replicate(targets, broadcast, src, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1 and broadcast:
        broadcast_targets(targets, src=src, group=group)


def coverage_sharded(grp1, start1, end1, grp2, start2, end2, n_groups, engine=None, group=None, broadcast=True, src=0):
    """This rank's slice of `coverage`: covered base pairs of its own query rows (ops.py:842-916).  Per-query
    results: the full answer is the concatenation of the shards in rank order, bit-exact."""
    if engine is None:
        from . import engine as engine
    _replicate([grp2, start2, end2], broadcast, src, group)
    return engine.coverage(grp1, start1, end1, grp2, start2, end2, n_groups)


def count_overlaps_sharded(grp1, start1, end1, grp2, start2, end2, n_groups, engine=None, group=None, broadcast=True,
                           src=0):
    """This rank's slice of `count_overlaps` (ops.py:1371-1438); concatenate the shards in rank order."""
    if engine is None:
        from . import engine as engine
    _replicate([grp2, start2, end2], broadcast, src, group)
    return engine.count_overlaps(grp1, start1, end1, grp2, start2, end2, n_groups)


def overlap_ordered_sharded(grp1, start1, end1, row_offset, grp2, start2, end2, n_groups, engine=None, group=None,
                            broadcast=True, src=0):
    """This rank's slice of `overlap(how='left', keep_order=True)` at the index level (ops.py:455-456, 549-550):
    rows in (query row, target row) order with GLOBAL query row numbers.  The result is per query, so the
    concatenation of the shards in rank order IS the single-GPU result, bit for bit, with the one broadcast
    (first bullet).  Returns (events1, events2)."""
    if engine is None:
        from . import engine as engine
    _replicate([grp2, start2, end2], broadcast, src, group)
    ev1, ev2, _ = engine.overlap_ordered(grp1, start1, end1, grp2, start2, end2, n_groups)
    ev1 = torch.as_tensor(ev1)
    if int(row_offset):
        ev1 = ev1 + int(row_offset)  # every row carries a query (never -1)
    return ev1, torch.as_tensor(ev2)


def subtract_sharded(grp1, start1, end1, row_offset, grp2, start2, end2, n_groups, engine=None, group=None,
                     broadcast=True, src=0):
    """This rank's slice of `subtract` (ops.py:1243-1330): the pieces of its own query rows, rows numbered
    globally; concatenate the shards in rank order.  Returns (row, start, end, sub_index)."""
    if engine is None:
        from . import engine as engine
    _replicate([grp2, start2, end2], broadcast, src, group)
    row, ps, pe, sub = engine.subtract(grp1, start1, end1, grp2, start2, end2, n_groups)
    row = torch.as_tensor(row)
    if int(row_offset):
        row = row + int(row_offset)
    return row, torch.as_tensor(ps), torch.as_tensor(pe), torch.as_tensor(sub)


def closest_sharded(grp1, start1, end1, row_offset, grp2, start2, end2, n_groups, engine=None, group=None,
                    broadcast=True, src=0, **kw):
    """This rank's slice of `_closest_intidxs` (ops.py:919-1040) with GLOBAL query row numbers.  Every query's
    rows depend only on the replicated targets.  Order of the concatenation: rank-major, inside a rank per group
    the matched rows then the unmatched ones; sorting the concatenated rows by (group block, matched-before-
    unmatched, query row) -- or simply by query row for frames that are assembled per query -- gives the
    single-GPU order.  Returns (events1, events2, n_matched_rows_of_this_shard)."""
    if engine is None:
        from . import engine as engine
    _replicate([grp2, start2, end2], broadcast, src, group)
    ev1, ev2, n_matched = engine.closest_intidxs(grp1, start1, end1, grp2, start2, end2, n_groups, **kw)
    ev1 = torch.as_tensor(ev1)
    if int(row_offset):
        ev1 = ev1 + int(row_offset)  # every row carries a query (never -1)
    return ev1, torch.as_tensor(ev2), n_matched
