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
closest per query).  `overlap_sharded` (row-range shards of unsorted queries) yields the pair SET in
(shard, group, A-block, B-block) order.

`overlap_range_sharded` is the exact form: the shards are contiguous (group, start) RANGES of the queries
(what the north star asks for), every rank gets the queries of its range plus the few halo queries a target
of its range can still reach, the targets are sorted ONCE on `src` and their sorted keys + row ids are
broadcast (12 B/row) while every rank sorts its own queries, and each rank produces only the rows it owns.
Per group the reference's rows (arrops.py:357-368, ops.py:319-326) then are: the A segments of the shards in
shard order, their B segments, their unpaired segments -- `assemble_offsets` turns the per-(shard, group)
segment sizes into the position of every segment in the reference's result, bit for bit."""
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


# ---------------------------------------------------------------------------------------------------------
# exact reference row order with (group, start)-range shards


def merge_extent_words(words, device=None, group=None):
    """all_reduce(MAX) of the mergeable extent record (bf_extent_measure) over the ranks -> numpy int64."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return np.asarray(words, dtype=np.int64)
    t = torch.from_numpy(np.ascontiguousarray(words, dtype=np.int64))
    if device is not None and torch.device(device).type == "cuda":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy()


def halo_reach(grp2, start2, end2, own_hi, n_groups):
    """How far beyond the upper bound `own_hi` = (g_hi, start_hi) of a shard its halo has to reach: the
    largest end' of a target that starts below the bound in the bound's own group (a target never leaves its
    group).  Queries of group g_hi with start in [start_hi, reach) are the shard's halo rows.  numpy, host."""
    g_hi, s_hi = int(own_hi[0]), int(own_hi[1])
    if g_hi >= n_groups:
        return s_hi
    g2, s2, e2 = np.asarray(grp2), np.asarray(start2), np.asarray(end2)
    m = (g2 == g_hi) & (s2 < s_hi) if n_groups > 1 else (s2 < s_hi)
    if not m.any():
        return s_hi
    e = e2[m] + (e2[m] == s2[m])
    return max(int(e.max()), s_hi)


def overlap_range_sharded(grp1, start1, end1, n_own, own_range, grp2, start2, end2, n_groups, how="left",
                          row_offset=0, halo_ids=None, engine=None, group=None, src=0, n2=None, device=None,
                          collective=True):
    """This rank's part of `_overlap_intidxs(df1, df2, how)` in the reference's row order.

    grp1/start1/end1: the queries of this rank's (group, start) range `own_range` = (g_lo, start_lo, g_hi,
    start_hi) -- rows [0, n_own), global row numbers row_offset + i -- followed by its halo rows (global row
    numbers `halo_ids`).  grp2/start2/end2: the targets, needed on `src` only (pass None elsewhere, with n2).
    Returns (events1, events2, n_pairs, counts[n_groups, 4]); `assemble_offsets(all ranks' counts)` places the
    per-group segments of events1/events2 in the whole result.  collective=False: no communication at all -- the
    caller holds the targets itself (one process walking over several shards, or a check)."""
    if engine is None:
        from . import engine as engine
    if how not in ("inner", "left"):
        raise ValueError("query-sharded overlap supports how='inner' and how='left' (unpaired targets need a reduction)")
    world = dist.get_world_size(group) if (dist.is_initialized() and collective) else 1
    rank = dist.get_rank(group) if (dist.is_initialized() and collective) else 0
    have_targets = world == 1 or rank == src
    use_cuda = torch.is_tensor(start1) and start1.is_cuda
    dev = start1.device if use_cuda else device

    import os
    import time

    trace = [] if os.environ.get("BF_PHASE_TIMES") and use_cuda else None

    def mark(name):
        if trace is not None:
            torch.cuda.synchronize(dev)
            trace.append((name, time.perf_counter()))

    mark("begin")
    # 1. one key layout for every set on every rank: local extents, MAX-merged (a few hundred bytes)
    words, presorted1 = engine.measure_extent(grp1, start1, end1, n_groups)
    presorted2 = False
    if have_targets:
        w2, presorted2 = engine.measure_extent(grp2, start2, end2, n_groups)
        words = np.maximum(words, w2)
        n2 = int(len(start2))
    mark("extent")
    if world > 1:
        words = merge_extent_words(words, device=dev, group=group)
    mark("allreduce")

    # 2. queries first: their sort is the long part (ms of GPU work), and once it is enqueued the host has all
    #    that time to enqueue the many small kernels of the target side behind it on a side stream.  Targets:
    #    sorted once on `src`, keys + row ids broadcast on the side stream while the queries sort.
    side = torch.cuda.Stream(device=dev) if use_cuda else None
    main = torch.cuda.current_stream(dev) if use_cuda else None
    if side is not None:  # the side stream may start once the inputs exist, not only after the query sort
        inputs_ready = torch.cuda.Event()
        inputs_ready.record(main)
        side.wait_event(inputs_ready)
    set1 = engine.prepare_set(words, grp1, start1, end1, n_groups, presorted=presorted1, slot="sharded.queries")
    mark("queries_sorted")
    ctx = torch.cuda.stream(side) if side is not None else _NullContext()
    handles = []
    with ctx:
        if have_targets:
            set2 = engine.prepare_set(words, grp2, start2, end2, n_groups, presorted=presorted2, slot="sharded.targets")
        else:
            set2 = engine.adopt_set(words, n2, n_groups, slot="sharded.targets")
        if world > 1:
            handles.append(dist.broadcast(set2.keys, src=src, group=group, async_op=True))
            handles.append(dist.broadcast(set2.ids, src=src, group=group, async_op=True))
        for h in handles:
            h.wait()  # the SIDE stream waits for the sorted targets to arrive; the main stream keeps sorting queries
        if how == "left" and hasattr(engine, "set_runmax"):
            engine.set_runmax(set2)  # the unpaired-row test reads the running maximum of the targets' ends: once per set
    if side is not None:
        main.wait_stream(side)  # the target store is persistent (SetStores): the streams are ordered here and by
        #                         `inputs_ready` of the next call, the allocator is not involved
    mark("targets_ready")

    # 3. search + count + emit, only what this shard owns
    sharded_call = world > 1 or own_range is not None
    ev1, ev2, n_pairs, counts = engine.overlap_prepared(
        set1, set2, how=how, own_range=own_range if sharded_call else None, row_offset1=int(row_offset),
        n_own1=int(n_own), halo_ids=halo_ids, want_group_counts=True)
    mark("count+emit")
    if trace is not None:
        print("[phases ms] " + "  ".join(f"{b[0]}={1e3 * (b[1] - a[1]):.2f}" for a, b in zip(trace, trace[1:])), flush=True)
    return ev1, ev2, n_pairs, counts


def _first_row_at_or_after(grp, start, g, s):
    lo, hi = np.searchsorted(grp, g, "left"), np.searchsorted(grp, g, "right")
    return int(lo + np.searchsorted(start[lo:hi], s, "left"))


def plan_range_shards(grp1, start1, world, grp2, start2, end2, n_groups):
    """Split queries that are sorted by (group rank, start) into `world` contiguous (group, start) ranges of
    about equal row counts, with the halo rows each shard needs.  Host side, numpy (the partition of df1 is the
    host's job).  -> list of dict(lo, hi: owned rows [lo, hi); halo: row numbers of the halo
    rows; own_range: (g_lo, start_lo, g_hi, start_hi))."""
    g1 = np.zeros(len(start1), np.int64) if grp1 is None else np.asarray(grp1)
    s1 = np.asarray(start1)
    n = len(s1)
    if n > 1 and not bool(np.all((g1[1:] > g1[:-1]) | ((g1[1:] == g1[:-1]) & (s1[1:] >= s1[:-1])))):
        raise ValueError("range shards need queries sorted by (group, start): their rows must be contiguous per shard")
    int_min = np.iinfo(np.int64).min
    cuts = [(0, int_min)]
    for r in range(1, world):
        p = min(r * n // world, max(n - 1, 0))
        cuts.append((int(g1[p]), int(s1[p])) if n else (n_groups, int_min))
    cuts.append((n_groups, int_min))
    shards = []
    for r in range(world):
        (g_lo, s_lo), (g_hi, s_hi) = cuts[r], cuts[r + 1]
        lo = _first_row_at_or_after(g1, s1, g_lo, s_lo) if r else 0
        hi = _first_row_at_or_after(g1, s1, g_hi, s_hi) if r + 1 < world else n
        hi = max(hi, lo)
        reach = halo_reach(grp2, start2, end2, (g_hi, s_hi), n_groups)
        h_hi = _first_row_at_or_after(g1, s1, g_hi, reach) if g_hi < n_groups else hi
        shards.append(dict(lo=lo, hi=hi, halo=np.arange(hi, max(h_hi, hi), dtype=np.int64),
                           own_range=(g_lo, s_lo, g_hi, s_hi)))
    return shards


class _NullContext:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def gather_group_counts(counts, device=None, group=None):
    """all_gather of the per-rank (n_groups, 4) segment sizes -> numpy [world, n_groups, 4]."""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return counts[None]
    t = torch.from_numpy(counts)
    if device is not None and torch.device(device).type == "cuda":
        t = t.to(device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy()


def assemble_offsets(counts_all):
    """counts_all [world, n_groups, 4] (A, B, unpaired 1, unpaired 2 rows per shard and group) ->
    (dst [world, n_groups, 3], src [world, n_groups, 3], total): rows src[r, c, t] ... + counts of shard r's
    result are rows dst[r, c, t] ... of the whole result, t = A / B / unpaired-1 segment of group c.  The
    reference's order per group: all A pairs (shards in (group, start) order), all B pairs, the unpaired rows
    (arrops.py:357-368, ops.py:319-326)."""
    c = np.asarray(counts_all, dtype=np.int64)[:, :, :3]
    world, nb, _ = c.shape
    # destination: group-major, then segment kind, then shard
    per = c.transpose(1, 2, 0).reshape(-1)  # [nb, 3, world]
    dst = (np.cumsum(per) - per).reshape(nb, 3, world).transpose(2, 0, 1)
    # source: inside a shard the rows are group-major, then A, B, unpaired (bf_overlap_emit)
    flat = c.reshape(world, -1)
    src = (np.cumsum(flat, axis=1) - flat).reshape(world, nb, 3)
    return dst, src, int(c.sum())


def assemble_host(parts, counts_all):
    """Concatenate-by-offset on the host: parts[r] = (events1, events2) numpy arrays of shard r -> the whole
    result in the reference's row order."""
    dst, src, total = assemble_offsets(counts_all)
    c = np.asarray(counts_all, dtype=np.int64)
    out1 = np.empty(total, dtype=np.int64)
    out2 = np.empty(total, dtype=np.int64)
    for r, (a, b) in enumerate(parts):
        for g in range(c.shape[1]):
            for t in range(3):
                k = int(c[r, g, t])
                if k:
                    out1[dst[r, g, t]:dst[r, g, t] + k] = a[src[r, g, t]:src[r, g, t] + k]
                    out2[dst[r, g, t]:dst[r, g, t] + k] = b[src[r, g, t]:src[r, g, t] + k]
    return out1, out2


# ---------------------------------------------------------------------------------------------------------
# frame level


def overlap_intidxs_frames(df1, df2, how="left", cols1=None, cols2=None, on=None, group=None, engine=None, gather=True):
    """`bioframe.ops._overlap_intidxs(df1, df2, how)` (ops.py:242-338) over the ranks of a process group: every
    rank holds the same two frames (df2 is only read on rank 0: its sorted keys reach the others by broadcast),
    takes the queries of its (chrom, start) range + halo, and -- with gather=True -- every rank returns the whole
    result in the reference's row order (events1, events2 as int64 numpy arrays); with gather=False only this
    rank's rows, their per-group segment sizes and the sizes of all ranks, for a caller that keeps the result
    distributed.  how='inner' / 'left'.

    The frames must be the same on every rank, and inside every (chrom, *on) group the starts of df1 must ascend in
    row order -- any frame sorted by chromosome and start, whatever the chromosome order -- because the unpaired rows
    of a group come in row order (ops.py:319-322) and shard order has to be row order for that (ValueError
    otherwise).  The visiting order of the groups follows the string hash seed of the process (set.union,
    ops.py:289): rank 0 fixes it for everyone."""
    from . import ops

    if engine is None:
        from . import engine as engine
    # all-null rows under a categorical chromosome column are in no group of the reference (ops._ungrouped_rows)
    gone1 = ops._ungrouped_rows(df1, (cols1 or ops._get_default_colnames())[0], on)
    gone2 = ops._ungrouped_rows(df2, (cols2 or ops._get_default_colnames())[0], on)
    if gone1 is not None or gone2 is not None:
        rows1 = None if gone1 is None else np.flatnonzero(~gone1)
        rows2 = None if gone2 is None else np.flatnonzero(~gone2)
        res = overlap_intidxs_frames(df1 if rows1 is None else df1.iloc[rows1], df2 if rows2 is None else df2.iloc[rows2],
                                     how=how, cols1=cols1, cols2=cols2, on=on, group=group, engine=engine, gather=gather)
        return (ops._frame_positions(res[0], rows1), ops._frame_positions(res[1], rows2), *res[2:])
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # group ranks on the host (the shard planner reads them); rank 0's visiting order for everyone
    enc = ops._encode_frames(df1, df2, cols1, cols2, on, _device=False)
    if world > 1:
        box = [enc[7] if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        if rank != 0:
            enc = ops._encode_frames(df1, df2, cols1, cols2, on, _group_order=box[0], _device=False)
    g1, s1, e1, g2, s2, e2, n_groups, _ = enc
    # the queries in (group rank, start) order: a stable sort by rank is enough when the starts of every group
    # already ascend in row order (plan_range_shards checks it)
    perm = np.argsort(g1, kind="stable")
    shards = plan_range_shards(g1[perm], s1[perm], world, g2, s2, e2, n_groups)
    sh = shards[rank]
    n_own = sh["hi"] - sh["lo"]
    rows = perm[np.concatenate([np.arange(sh["lo"], sh["hi"]), sh["halo"]])]  # frame rows of this shard, halo last
    use_cuda = torch.cuda.is_available() and getattr(engine, "__name__", "").endswith("engine") and hasattr(engine, "Workspace")
    put = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if use_cuda else (lambda a: a)
    targets = (put(g2), put(s2), put(e2)) if rank == 0 else (None, None, None)
    ev1, ev2, _n_pairs, counts = overlap_range_sharded(
        put(g1[rows]), put(s1[rows]), put(e1[rows]), n_own, sh["own_range"], *targets, n_groups, how=how,
        halo_ids=put(np.arange(n_own, len(rows), dtype=np.int64)), engine=engine, group=group, n2=len(s2))
    host = lambda t: t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)  # noqa: E731
    ev1, ev2 = host(ev1), host(ev2)
    ev1 = np.where(ev1 >= 0, rows[np.maximum(ev1, 0)], -1) if len(rows) else ev1  # shard-local -> frame row numbers
    counts_all = gather_group_counts(counts, device=("cuda" if use_cuda else None), group=group)
    if not gather:
        return ev1, ev2, counts, counts_all
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (ev1, ev2), group=group)
    else:
        parts = [(ev1, ev2)]
    return assemble_host(parts, counts_all)
