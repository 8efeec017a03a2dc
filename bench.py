#!/usr/bin/env python
"""bench.py -- the north-star benchmark: `bf.overlap` index path at 1e8 x 1e7.

One "step" = one pass of `_overlap_intidxs(how='left')` over
the cfg3 workload of BASELINE.json: 1e8 synthetic hg38 queries (Geometric mean
2 kb) against 1e7 peak-like targets, ~9 overlaps per query, ~9e8 output rows.

  value        (N1*n_gpus + N2) / t, inputs resident in HBM, CUDA-event timed
  e2e          same call with HOST (pinned) buffers: H2D of the three columns of
               both sets and D2H of both int64 result columns inside the timed region
  roofline     dominant kernel = one onesweep radix pass over the query keys
  cpu_baseline the REFERENCE itself (`bioframe.ops._overlap_intidxs` from the vendored
               baseline/_ref, oracle/make_ref.py; the numpy port only if that copy is
               missing) on a bounded, same-density sample (chr21+chr22), 1 core
  parity_checked  the GPU rows of that sample equal the reference's rows, in order
  frame_e2e    DataFrames in, DataFrames out through `bioframe_b200.overlap` /
               `ops._overlap_intidxs`, the reference's `bioframe.overlap` beside it,
               and where the call spends its time (encode / H2D / kernels / D2H / frame)
  secondary    the other BASELINE.json configs (cfg0 overlap inner 1e6 x 1e6, cfg1
               closest, cfg2 cluster, cfg3 coverage) at 3 steps each

`--impl reference` times the reference's `_overlap_intidxs` on DataFrames with a
categorical chromosome column, all 24 chromosomes of the workload at full density
spread over every host core (one process per core, chromosomes bin-packed).

N > 1 (torchrun): the queries are sharded by contiguous (chrom, start) ranges (rank r
owns the r-th slice of the end-to-end genome axis and 1e8 queries in it: weak
scaling; `--scaling strong`: 1e8 queries in total), the targets are sorted once on
rank 0 and their sorted keys + row ids are broadcast over NCCL on a side stream
while every rank sorts its own queries; the only other collective is one MAX
all-reduce of a few hundred bytes (the shared key layout).  Every rank produces its
part of the reference's row order (bioframe_b200/sharded.py); after the timed region
the result of every rank is checked against the single-GPU engine
(`sharded_verified`).  `--workload cfg4`: BASELINE.json configs[4], 1e9 sorted
queries x 1e8 targets over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

_T_START = time.perf_counter()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GROUPS = 24
SAMPLE_CHROMS = (20, 21)  # chr21 + chr22: the bounded CPU sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n1", type=int, default=100_000_000)
    ap.add_argument("--n2", type=int, default=10_000_000)
    ap.add_argument("--how", default="left", choices=["inner", "left", "right", "outer"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--chunks", type=int, default=4, help="cfg4: query chunks (virtual shards) per rank")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-frame", action="store_true", help="skip the DataFrame-level leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE.json configs")
    ap.add_argument("--time-budget", type=float, default=240.0,
                    help="seconds since process start after which the reported extra legs (all-core reference, frames, "
                         "secondary configs) are skipped, so that a slow host cannot cost the headline line")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the check against the single-GPU engine")
    ap.add_argument("--force-ranged", action="store_true", help="run the range-sharded path even on one GPU (diagnostic)")
    ap.add_argument("--workload", default="overlap",
                    choices=["overlap", "cfg4", "cfg0", "closest", "cluster", "coverage", "count", "subtract", "ingest", "ordered"],
                    help="overlap = the north-star line (cfg3); cfg4 = 1e9 x 1e8 over the ranks; the others are the "
                         "secondary configs (cfg0: overlap inner 1e6 x 1e6; closest: cfg1 1e7 x 1e6 k=1; cluster: cfg2 1e8 "
                         "heavy-tailed; coverage: cfg3 sets)")
    ap.add_argument("--k", type=int, default=1)
    return ap.parse_args()


_WL_CACHE = {}


def workload(n1, n2, rank, chrom_ids=None):
    """cfg3 sets (host numpy); the full-size draw is kept for the legs of one run that share it."""
    from bioframe_b200 import synth

    key = (n1, n2, rank, tuple(chrom_ids) if chrom_ids is not None else None)
    if key in _WL_CACHE:
        return _WL_CACHE[key]
    frac = synth.subset_fraction(chrom_ids) if chrom_ids is not None else 1.0
    m1 = max(int(n1 * frac), 1)
    q = synth.make_intervals(m1, 1 + 1000 * rank, "geom2k", chrom_ids=chrom_ids)
    t = targets(n2, chrom_ids)
    if chrom_ids is None:
        _WL_CACHE.clear()
        _WL_CACHE[key] = (q, t)
    return q, t


def targets(n2, chrom_ids=None):
    from bioframe_b200 import synth

    frac = synth.subset_fraction(chrom_ids) if chrom_ids is not None else 1.0
    hot = max(int(200_000 * frac * (n2 / 1e7)), 1)
    return synth.make_peak_targets(max(int(n2 * frac), 1), 2, n_hotspots=hot, chrom_ids=chrom_ids)


def workload_desc(how, n1, n2):
    return (f"_overlap_intidxs(how={how!r}) cfg3: {n1} hg38 queries (geom 2kb) per GPU x {n2} peak-like targets, "
            "unsorted")


def cores_available():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ reference / oracle
def reference_module():
    """The vendored reference (baseline/_ref, oracle/make_ref.py) or None.  /root/reference is never read."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import make_ref

        return make_ref.import_reference()
    except Exception:
        return None


def reference_frame(g, s, e, categorical=True):
    """hg38 interval frame the way a bioframe user holds it: chrom as names (categorical: the fast case of
    docs/guide-performance.ipynb:497-498)."""
    import pandas as pd

    from bioframe_b200 import synth

    if categorical:
        chrom = pd.Categorical.from_codes(np.asarray(g), categories=synth.HG38_CHROMS)
    else:
        chrom = np.array(synth.HG38_CHROMS, dtype=object)[np.asarray(g)]
    return pd.DataFrame({"chrom": chrom, "start": s, "end": e})


def cpu_overlap(q, t, how, order_by_code=False):
    """One pass of the reference path on the host.  -> (events1, events2, kind).  With order_by_code the chromosome
    column holds the integer group codes, which makes the reference visit its groups in code order (the C ABI's
    order) instead of the process's string-hash order: used where the rows are compared."""
    bf = reference_module()
    if bf is not None:
        import pandas as pd

        if order_by_code:
            df1 = pd.DataFrame({"chrom": q[0].astype(np.int64), "start": q[1], "end": q[2]})
            df2 = pd.DataFrame({"chrom": t[0].astype(np.int64), "start": t[1], "end": t[2]})
        else:
            df1, df2 = reference_frame(*q), reference_frame(*t)
        a, b = bf.ops._overlap_intidxs(df1, df2, how=how)
        return np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64), "reference"
    from oracle import ops_oracle as oo

    a, b = oo.overlap_intidxs_coded(*q, *t, N_GROUPS, how=how)
    return a, b, "port"


def cpu_leg(args, steps, warmup):
    """The reference on the bounded sample, 1 core; returns a dict with the rate, the sample and its rows."""
    q, t = workload(args.n1, args.n2, 0, chrom_ids=SAMPLE_CHROMS)
    times, rows, kind = [], 0, "port"
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ev1, ev2, kind = cpu_overlap(q, t, args.how, order_by_code=True)
        dt = time.perf_counter() - t0
        rows = int(ev1.shape[0])
        if it >= warmup:
            times.append(dt)
    tm = float(np.mean(times))
    n = len(q[1]) + len(t[1])
    what = "bioframe.ops._overlap_intidxs (reference v0.8.0, baseline/_ref)" if kind == "reference" else "numpy oracle port of ops._overlap_intidxs"
    desc = (f"chr21+chr22 share of the workload at full density: {len(q[1])} x {len(t[1])} intervals -> {rows} rows per step, "
            f"{what}(how={args.how!r})")
    return dict(value=n / tm, desc=desc, ms=tm * 1e3, rows=rows, kind=kind, q=q, t=t, ev1=ev1, ev2=ev2)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as exc:  # measurement aid only
            self.reasons.add(f"sampler_error:{type(exc).__name__}")

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pin_to_gpu_numa_node(index):
    """Run this rank on the CPU cores next to its GPU, so that the pinned host buffers of the e2e leg
    (allocated first-touch) sit in that NUMA node's memory: with 8 ranks on one box the host copies
    otherwise pile up on one memory controller.  Best effort; a no-op when NVML gives no affinity."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ------------------------------------------------------------------------------------------ --impl reference
def _reference_worker(conn, n1, n2, how, chroms):
    """One process of the reference arm: the reference's `_overlap_intidxs` (ops.py:242-338) on DataFrames that
    hold this worker's chromosomes of the workload at full density, run whenever the parent says go."""
    q, t = workload(n1, n2, 0, chrom_ids=tuple(chroms))
    bf = reference_module()
    if bf is not None:
        df1, df2 = reference_frame(*q), reference_frame(*t)
        call = lambda: bf.ops._overlap_intidxs(df1, df2, how=how)  # noqa: E731
        kind = "reference"
    else:
        from oracle import ops_oracle as oo

        call = lambda: oo.overlap_intidxs_coded(*q, *t, N_GROUPS, how=how)  # noqa: E731
        kind = "port"
    conn.send((len(q[1]), len(t[1]), kind))
    while conn.recv():
        ev1, _ = call()
        conn.send(int(len(ev1)))


def pack_chromosomes(n_bins):
    """All 24 chromosomes over n_bins workers, longest first into the lightest bin (the work is ~linear in length)."""
    from bioframe_b200 import synth

    bins = [[] for _ in range(n_bins)]
    load = [0] * n_bins
    for c in np.argsort(-synth.HG38_SIZES):
        b = int(np.argmin(load))
        bins[b].append(int(c))
        load[b] += int(synth.HG38_SIZES[c])
    return [sorted(b) for b in bins if b]


def reference_leg_all_cores(args, steps, warmup):
    """The reference on every host core: its chromosome groups are independent (ops.py:301-338), so one process
    per core, each holding a share of the chromosomes, is how its user spreads it over a machine.  ALL 24
    chromosomes of the workload at full density; a step ends when the slowest process is done.
    Returns (intervals/s, description, ms/step, rows, cores, kind)."""
    import multiprocessing as mp

    cores = max(1, min(cores_available(), N_GROUPS))
    bins = pack_chromosomes(cores)
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
    procs = []
    for chroms in bins:
        parent, child = ctx.Pipe()
        pr = ctx.Process(target=_reference_worker, args=(child, args.n1, args.n2, args.how, chroms), daemon=True)
        pr.start()
        procs.append((pr, parent))
    sizes = [conn.recv() for _, conn in procs]
    kind = sizes[0][2]
    times, rows = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for _, conn in procs:
            conn.send(True)
        rows = sum(conn.recv() for _, conn in procs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    for pr, conn in procs:
        conn.send(False)
        pr.join(timeout=10)
    tm = float(np.mean(times))
    n = sum(a + b for a, b, _ in sizes)
    what = ("bioframe.ops._overlap_intidxs of the reference (v0.8.0, baseline/_ref) on DataFrames with a categorical chrom column"
            if kind == "reference" else "numpy oracle port of ops._overlap_intidxs (baseline/_ref missing)")
    desc = (f"{len(bins)} processes sharing all 24 chromosomes of the workload at full density: "
            f"{sum(a for a, _, _ in sizes)} x {sum(b for _, b, _ in sizes)} intervals -> {rows} rows per step, {what}, how={args.how!r}")
    return n / tm, desc, tm * 1e3, rows, len(bins), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, desc, ms, rows, cores, kind = reference_leg_all_cores(args, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": "bf.overlap input intervals/s",
        "value": val,
        "unit": "intervals/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": workload_desc(args.how, args.n1, args.n2), "n1_per_gpu": args.n1, "n2": args.n2,
                   "rows_out": rows, "sample": desc},
        "cpu_baseline": {"value": val, "unit": "intervals/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ secondary configs
def _ingest_device(text, raw, dev):
    """The device part of bioframe_b200.io.read_bed_device on text that already is on the GPU; returns the row count."""
    import ctypes as C

    import torch

    from bioframe_b200 import _lib
    from bioframe_b200.engine import Workspace, _ptr, _stream_ptr

    L = _lib.lib()
    n_bytes = int(text.numel())
    stream = _stream_ptr(dev)
    ws = Workspace.get(dev, max(L.bf_bed_scan_workspace_bytes(n_bytes), 256))
    n_lines = C.c_int64(0)
    _lib.check(L.bf_bed_scan(_ptr(text), n_bytes, _ptr(ws), ws.numel(), stream, C.byref(n_lines)), "bf_bed_scan")
    ws = Workspace.get(dev, max(L.bf_bed_parse_workspace_bytes(n_bytes, n_lines.value), 256))
    plan = _lib.Plan()
    n_rows, n_names = C.c_int64(0), C.c_int32(0)
    first, length, slot = np.zeros(1 << 15, np.int64), np.zeros(1 << 15, np.int32), np.zeros(1 << 15, np.int32)
    _lib.check(L.bf_bed_parse(_ptr(text), n_bytes, n_lines.value, _ptr(ws), ws.numel(), C.byref(plan), stream, C.byref(n_rows),
                              C.byref(n_names), C.c_void_p(first.ctypes.data), C.c_void_p(length.ctypes.data),
                              C.c_void_p(slot.ctypes.data)), "bf_bed_parse")
    k = n_names.value
    found = [bytes(raw[first[i]: first[i] + length[i]]) for i in range(k)]
    slot_code = np.full(1 << 16, -1, dtype=np.int32)
    for code, i in enumerate(sorted(range(k), key=lambda i: found[i])):
        slot_code[slot[i]] = code
    chrom = torch.empty(n_rows.value, dtype=torch.int32, device=dev)
    start = torch.empty(n_rows.value, dtype=torch.int64, device=dev)
    end = torch.empty(n_rows.value, dtype=torch.int64, device=dev)
    _lib.check(L.bf_bed_emit(C.byref(plan), _ptr(ws), C.c_void_p(slot_code.ctypes.data), _ptr(chrom), _ptr(start), _ptr(end),
                             stream), "bf_bed_emit")
    return n_rows.value


def _named_frame(g, s, e):
    """frame with string chromosome names whose sorted order is the code order (what cluster / merge / coverage of
    the reference need: they validate the chrom dtype, ops.py:758)"""
    import pandas as pd

    names = np.array([f"c{c:02d}" for c in range(N_GROUPS)], dtype=object)
    return pd.DataFrame({"chrom": names[np.asarray(g)], "start": s, "end": e})


def secondary_line(args, wl, steps, warmup, with_cpu=True):
    """One secondary config: device-resident inputs, CUDA-event timing of the whole call, whole-call algorithmic
    bytes against the measured HBM peak, and the reference on a bounded sample beside it."""
    import torch

    from bioframe_b200 import _lib, engine, synth
    from oracle import ops_oracle as oo

    dev = torch.device("cuda", torch.cuda.current_device())
    bf = reference_module() if with_cpu else None
    cpu_kind = "reference" if bf is not None else "port"
    cpu_sample = None
    frac = synth.subset_fraction(SAMPLE_CHROMS)
    if wl == "cfg0":
        n1, n2 = 1_000_000, 1_000_000
        q, t = synth.make_intervals(n1, 1, "geom1k"), synth.make_intervals(n2, 2, "geom1k")
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.overlap_intidxs(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS, how="inner")  # noqa: E731
        rows_of = lambda r: int(r[0].numel())  # noqa: E731
        a_alg = lambda R: 256 * (n1 + n2) + 20 * R  # noqa: E731
        units, name = n1 + n2, "bf.overlap(how='inner') input intervals/s"
        desc = f"_overlap_intidxs(how='inner') cfg0: {n1} x {n2} hg38 intervals (geom 1kb), unsorted"
        cpu = lambda: cpu_overlap(q, t, "inner")  # noqa: E731  (the WHOLE config: it is the reference's own CPU-runnable case)
        cpu_units = n1 + n2
        cpu_sample = f"the whole cfg0 workload ({n1} x {n2}), bioframe.ops._overlap_intidxs on categorical-chrom DataFrames" if bf is not None else None
    elif wl == "closest":
        n1, n2 = (args.n1 if args.n1 != 100_000_000 else 10_000_000), (args.n2 if args.n2 != 10_000_000 else 1_000_000)
        q = synth.make_intervals(n1, 1, "geom1k")
        t = synth.make_intervals(n2, 2, "geom1k")
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.closest_intidxs(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS, k=args.k)  # noqa: E731
        rows_of = lambda r: int(r[0].numel())  # noqa: E731
        a_alg = lambda R: 2 * 256 * n2 + 36 * n1 + 20 * R  # noqa: E731
        units, name = n1 + n2, f"bf.closest(k={args.k}) input intervals/s"
        desc = f"_closest_intidxs cfg1: {n1} hg38 queries x {n2} targets (geom 1kb), k={args.k}, unsorted"
        sq = synth.make_intervals(int(n1 * frac), 1, "geom1k", chrom_ids=SAMPLE_CHROMS)
        stt = synth.make_intervals(int(n2 * frac), 2, "geom1k", chrom_ids=SAMPLE_CHROMS)
        if bf is not None:
            f1, f2 = reference_frame(*sq), reference_frame(*stt)
            cpu = lambda: bf.ops._closest_intidxs(f1, f2, k=args.k)  # noqa: E731
        else:
            cpu = lambda: oo.closest_intidxs_coded(*sq, *stt, N_GROUPS, k=args.k)  # noqa: E731
        cpu_units = len(sq[1]) + len(stt[1])
    elif wl == "cluster":
        n1 = args.n1
        q = synth.make_intervals(n1, 1, "pareto")
        d = [torch.from_numpy(a).to(dev) for a in q]
        call = lambda: engine.cluster(d[0], d[1], d[2], N_GROUPS, min_dist=0, want_ids=True, want_table=True)  # noqa: E731
        rows_of = lambda r: int(r["n_clusters"])  # noqa: E731
        a_alg = lambda C: 256 * n1 + 12 * n1 + 28 * C  # noqa: E731
        units, name = n1, "bf.cluster input intervals/s"
        desc = f"cluster+merge table cfg2: {n1} hg38 intervals, Pareto(1.2) lengths capped 5 Mb, unsorted"
        sq = synth.make_intervals(int(n1 * frac), 1, "pareto", chrom_ids=SAMPLE_CHROMS)
        if bf is not None:
            f1 = _named_frame(*sq)
            cpu = lambda: bf.cluster(f1, min_dist=0)  # noqa: E731
        else:
            cpu = lambda: oo.cluster_coded(*sq, N_GROUPS, 0)  # noqa: E731
        cpu_units = len(sq[1])
    elif wl == "count":
        n1, n2 = args.n1, args.n2
        (q, t) = workload(n1, n2, 0)
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.count_overlaps(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS)  # noqa: E731
        rows_of = lambda r: int(r.numel())  # noqa: E731
        a_alg = lambda R: 2 * 256 * n2 + 36 * n1  # noqa: E731
        units, name = n1 + n2, "bf.count_overlaps input intervals/s"
        desc = f"count_overlaps cfg3: {n1} hg38 queries (geom 2kb) x {n2} peak-like targets"
        (sq, stt) = workload(n1, n2, 0, chrom_ids=SAMPLE_CHROMS)
        if bf is not None:
            f1, f2 = _named_frame(*sq), _named_frame(*stt)
            cpu = lambda: bf.count_overlaps(f1, f2)  # noqa: E731
        else:
            cpu = lambda: np.bincount(oo.overlap_intidxs_coded(*sq, *stt, N_GROUPS, how="inner")[0], minlength=len(sq[1]))  # noqa: E731
        cpu_units = len(sq[1]) + len(stt[1])
    elif wl == "ingest":
        import io as _io

        import pandas as pd

        n1 = args.n1 if args.n1 != 100_000_000 else 10_000_000
        q = synth.make_intervals(n1, 1, "geom2k")
        names = np.array(synth.HG38_CHROMS, dtype=object)
        buf = _io.BytesIO()
        pd.DataFrame({"c": names[q[0]], "s": q[1], "e": q[2]}).to_csv(buf, sep="\t", header=False, index=False)
        raw = np.frombuffer(buf.getvalue(), dtype=np.uint8)
        text = torch.from_numpy(raw.copy()).to(dev)
        nbytes = int(raw.size)

        def call():
            # device-resident text -> device columns (the H2D of the text is the e2e part, not timed here)
            return _ingest_device(text, raw, dev)

        rows_of = lambda r: int(r)  # noqa: E731
        a_alg = lambda R: 3 * nbytes + 52 * R  # noqa: E731  (text read three times; line table, parsed fields, outputs)
        units, name = n1, "BED ingest rows/s"
        desc = f"bed3 ingest: {n1} lines, {nbytes} bytes of tab-separated text (hg38 names, geom 2kb intervals)"
        cpu = lambda: pd.read_csv(_io.BytesIO(buf.getvalue()), sep="\t", header=None, names=["chrom", "start", "end"])  # noqa: E731
        cpu_units = n1
        cpu_kind = "reference"
        cpu_sample = "the same text through pandas.read_csv(sep=TAB, header=None), which is what the reference's read_table calls"
    elif wl == "ordered":
        n1, n2 = args.n1, args.n2
        (q, t) = workload(n1, n2, 0)
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.overlap_ordered(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS)  # noqa: E731
        rows_of = lambda r: int(r[0].numel())  # noqa: E731
        a_alg = lambda R: 2 * 256 * n2 + 36 * n1 + 20 * R  # noqa: E731  (targets sorted twice, probes, pairs)
        units, name = n1 + n2, "bf.overlap(how='left', keep_order=True) input intervals/s"
        desc = f"left overlap in row order (keep_order) cfg3: {n1} hg38 queries (geom 2kb) x {n2} peak-like targets"
        (sq, stt) = workload(n1, n2, 0, chrom_ids=SAMPLE_CHROMS)

        def cpu():
            a, b, _ = cpu_overlap(sq, stt, "left")
            return np.lexsort([np.where(b < 0, len(stt[1]) + 1, b), a])

        cpu_units = len(sq[1]) + len(stt[1])
    elif wl == "subtract":
        n1, n2 = args.n1, args.n2
        (q, t) = workload(n1, n2, 0)
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.subtract(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS)  # noqa: E731
        rows_of = lambda r: int(r[0].numel())  # noqa: E731
        a_alg = lambda R: 256 * n2 + 36 * n1 + 32 * R  # noqa: E731  (merge of the targets, probes, 4 x 8 B per piece)
        units, name = n1 + n2, "bf.subtract input intervals/s"
        desc = f"subtract cfg3: {n1} hg38 queries (geom 2kb) minus {n2} peak-like targets"
        (sq, stt) = workload(n1, n2, 0, chrom_ids=SAMPLE_CHROMS)
        if bf is not None:
            f1, f2 = _named_frame(*sq), _named_frame(*stt)
            cpu = lambda: bf.subtract(f1, f2)  # noqa: E731
        else:
            cpu = lambda: oo.subtract_coded(*sq, *stt, N_GROUPS)  # noqa: E731
        cpu_units = len(sq[1]) + len(stt[1])
    else:  # coverage
        n1, n2 = args.n1, args.n2
        (q, t) = workload(n1, n2, 0)
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.coverage(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS)  # noqa: E731
        rows_of = lambda r: int(r.numel())  # noqa: E731
        a_alg = lambda R: 256 * n2 + 36 * n1  # noqa: E731
        units, name = n1 + n2, "bf.coverage input intervals/s"
        desc = f"coverage cfg3: {n1} hg38 queries (geom 2kb) x {n2} peak-like targets"
        (sq, stt) = workload(n1, n2, 0, chrom_ids=SAMPLE_CHROMS)
        if bf is not None:
            f1, f2 = _named_frame(*sq), _named_frame(*stt)
            cpu = lambda: bf.coverage(f1, f2)  # noqa: E731
        else:
            cpu = lambda: oo.coverage_coded(*sq, *stt, N_GROUPS)  # noqa: E731
        cpu_units = len(sq[1]) + len(stt[1])

    for _ in range(max(warmup, 3)):
        r = call()
    R = rows_of(r)
    del r
    _lib.prof_reset()
    _lib.prof_enable(True)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        r = call()
        del r
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1) / steps
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    peak, _src = hbm_peak()
    alg = a_alg(R)
    line = {
        "metric": name, "value": units / (ms * 1e-3), "unit": "intervals/s", "n_gpus": 1, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc, "rows_out": R, "l2": "inputs and intermediates exceed the 126 MB L2" if units > 5e6 else
                   "inputs fit the 126 MB L2: launch-latency bound (SURVEY.md 8d), reported, not tuned for"},
        "gpu_launches": sum(v["launches"] for v in prof.values()),
        "roofline": {"bound": "hbm", "kernel": "whole call (A_alg of SURVEY.md 8d)", "achieved": alg / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()}},
    }
    if with_cpu:
        times = []
        for it in range(2):
            w0 = time.perf_counter()
            cpu()
            times.append(time.perf_counter() - w0)
        what = "the reference (bioframe v0.8.0, baseline/_ref)" if cpu_kind == "reference" else "numpy oracle port"
        line["cpu_baseline"] = {"value": cpu_units / min(times), "unit": "intervals/s", "cores": 1, "kind": cpu_kind,
                                "sample": cpu_sample or f"chr21+chr22 share of the workload at full density ({cpu_units} intervals), {what}",
                                "ms_per_step": min(times) * 1e3}
    del d
    torch.cuda.empty_cache()
    return line


def run_secondary(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(0)
    print(json.dumps(secondary_line(args, args.workload, args.steps, args.warmup, with_cpu=not args.no_cpu)), flush=True)


# ------------------------------------------------------------------------------------------ DataFrame-level leg
def frame_leg(args, over_budget=lambda: False):
    """DataFrames in, DataFrames out (VERDICT r1 Missing #7): the public `bioframe_b200.overlap` and the seam
    `ops._overlap_intidxs` on pandas frames (categorical chrom + a string `name` column), wall clock around the
    whole call -- validation, group encoding, H2D, kernels, D2H, frame assembly -- with the reference's own call
    timed beside it on the same frames, and the equality of the two results."""
    import pandas as pd

    import bioframe_b200 as ours
    from bioframe_b200 import ops as our_ops
    from bioframe_b200 import synth

    bf = reference_module()
    out = {}

    def frames(n1, n2, kind1, kind2, categorical=True):
        q, t = synth.make_intervals(n1, 1, kind1), synth.make_intervals(n2, 2, kind2)
        df1, df2 = reference_frame(*q, categorical=categorical), reference_frame(*t, categorical=categorical)
        df1["name"] = np.array([f"q{i}" for i in range(1000)], dtype=object)[np.arange(n1) % 1000]
        df2["score"] = (np.arange(n2) % 97).astype(np.float64)
        return df1, df2

    def wall(fn, reps):
        best, res = None, None
        for _ in range(reps):
            w0 = time.perf_counter()
            res = fn()
            dt = time.perf_counter() - w0
            best = dt if best is None else min(best, dt)
        return best, res

    # SURVEY.md 8d asks for the chromosome column as names (object / str) AND as a categorical: the third entry is cfg0
    # with plain string names, where finding the groups is a hash pass over the rows on the host for both libraries
    for tag, n1, n2, categorical in (("cfg0_1e6x1e6", 1_000_000, 1_000_000, True), ("1e7x1e6", 10_000_000, 1_000_000, True),
                                     ("cfg0_1e6x1e6_string_chrom", 1_000_000, 1_000_000, False)):
        df1, df2 = frames(n1, n2, "geom1k", "geom1k", categorical)
        entry = {"n1": n1, "n2": n2, "columns": f"chrom ({'categorical' if categorical else str(df1['chrom'].dtype)}), start, end, "
                                                "name (str) | chrom, start, end, score (f64)"}
        ours.overlap(df1.iloc[:1000], df2.iloc[:1000], how="inner")  # warm the library
        t_idx, idx = wall(lambda: our_ops._overlap_intidxs(df1, df2, how="inner"), 2)
        t_all, got = wall(lambda: ours.overlap(df1, df2, how="inner"), 2)
        with our_ops.profile() as prof:
            ours.overlap(df1, df2, how="inner")
        stages = {k: round(v * 1e3, 2) for k, v in prof.items()}
        pandas_ms = stages.get("validate", 0) + stages.get("encode_groups", 0) + stages.get("assemble_frame", 0)
        entry["bioframe_b200"] = {"_overlap_intidxs_ms": t_idx * 1e3, "overlap_inner_ms": t_all * 1e3, "rows": int(len(got)),
                                  "intervals_per_s": (n1 + n2) / t_all, "stages_ms_profiled": stages,
                                  "pandas_share_profiled": pandas_ms / max(sum(stages.values()), 1e-9)}
        if bf is not None:
            r_idx, ridx = wall(lambda: bf.ops._overlap_intidxs(df1, df2, how="inner"), 1)
            r_all, want = wall(lambda: bf.overlap(df1, df2, how="inner"), 1)
            key = [c for c in got.columns if c not in ("chrom", "chrom_")]
            same_pairs = bool(np.array_equal(np.sort(idx[0] * (n2 + 1) + idx[1]), np.sort(ridx[0] * (n2 + 1) + ridx[1])))
            a = got.sort_values(key, kind="stable").reset_index(drop=True)
            b = want.sort_values(key, kind="stable").reset_index(drop=True)
            try:
                pd.testing.assert_frame_equal(a, b)
                same_frame = True
            except AssertionError:
                same_frame = False
            entry["reference"] = {"_overlap_intidxs_ms": r_idx * 1e3, "overlap_inner_ms": r_all * 1e3, "cores": 1,
                                  "intervals_per_s": (n1 + n2) / r_all}
            entry["speedup_overlap_inner"] = r_all / t_all
            entry["speedup_overlap_intidxs"] = r_idx / t_idx
            entry["frames_equal_to_reference"] = same_frame and same_pairs
            entry["note"] = "rows compared after sorting both frames: the reference's chromosome-block order follows the string hash seed"
        out[tag] = entry
        del df1, df2, got
    # the other public calls of BASELINE.json's configs at a size the reference finishes in seconds (its closest needs
    # 21 s at 1e7 x 1e6, its coverage 17 s: SURVEY.md 6): frames in, frames out, results compared
    if over_budget():
        out["skipped_after_overlap"] = "time budget"
        return out
    if bf is not None:
        n1, n2 = 3_000_000, 300_000
        q = synth.make_intervals(n1, 1, "geom1k")
        t = synth.make_tie_free(n2, 2, "geom1k")  # closest among equal ends / starts is undefined in the reference
        df1, df2 = reference_frame(*q), reference_frame(*t)
        heavy = reference_frame(*synth.make_intervals(n1, 3, "pareto"))
        other = {}

        def same(a, b, canon=False):
            try:
                if canon:
                    key = list(a.columns)
                    a = a.sort_values(key, kind="stable").reset_index(drop=True)
                    b = b.sort_values(key, kind="stable").reset_index(drop=True)
                pd.testing.assert_frame_equal(a, b)
                return True
            except AssertionError:
                return False

        for name, ours_fn, ref_fn, canon in (
            ("closest_k1", lambda: ours.closest(df1, df2, k=1), lambda: bf.closest(df1, df2, k=1), True),
            ("cluster", lambda: ours.cluster(heavy), lambda: bf.cluster(heavy), False),
            ("merge", lambda: ours.merge(heavy), lambda: bf.merge(heavy), False),
            ("coverage", lambda: ours.coverage(df1, df2), lambda: bf.coverage(df1, df2), False),
            ("count_overlaps", lambda: ours.count_overlaps(df1, df2), lambda: bf.count_overlaps(df1, df2), False),
            ("subtract", lambda: ours.subtract(df1, df2), lambda: bf.subtract(df1, df2), False),
        ):
            ours_fn()
            t_ours, got = wall(ours_fn, 2)
            t_ref, want = wall(ref_fn, 1)
            other[name] = {"bioframe_b200_ms": t_ours * 1e3, "reference_ms": t_ref * 1e3, "speedup": t_ref / t_ours,
                           "frames_equal_to_reference": same(got, want, canon)}
            del got, want
        out["other_ops_3e6x3e5"] = {"sizes": f"{n1} queries (geom 1kb; cluster / merge: Pareto lengths) x {n2} tie-free targets, "
                                             "categorical chrom", **other}
        del df1, df2, heavy
    if over_budget():
        out["skipped_after_other_ops"] = "time budget"
        return out
    # the seam at the benchmark size: DataFrames (categorical chrom) -> int64 numpy index arrays on the host
    q, t = workload(args.n1, args.n2, 0)
    df1, df2 = reference_frame(*q), reference_frame(*t)
    del q, t
    t_idx, idx = wall(lambda: our_ops._overlap_intidxs(df1, df2, how=args.how), 1)
    with our_ops.profile() as prof:
        our_ops._overlap_intidxs(df1, df2, how=args.how)
    out["cfg3_intidxs_frames"] = {
        "what": f"bioframe_b200.ops._overlap_intidxs(df1, df2, how={args.how!r}) on DataFrames of {len(df1)} / {len(df2)} rows",
        "ms": t_idx * 1e3, "rows": int(len(idx[0])), "intervals_per_s": (len(df1) + len(df2)) / t_idx,
        "stages_ms_profiled": {k: round(v * 1e3, 1) for k, v in prof.items()},
    }
    del idx
    # the public call at the benchmark size: the whole result FRAME (8.76e8 rows x 6 columns, ~40 GB on the host).  The
    # reference materialises every input column per pair through numpy and needs > 60 GB and minutes for it (SURVEY.md
    # 8d); only attempted where the host has the memory
    try:
        import psutil

        room = psutil.virtual_memory().available
    except Exception:
        room = 0
    if room > 150e9:
        w0 = time.perf_counter()
        res = ours.overlap(df1, df2, how="inner")
        dt = time.perf_counter() - w0
        out["cfg3_overlap_inner_frames"] = {
            "what": f"bioframe_b200.overlap(df1, df2, how='inner') on DataFrames of {len(df1)} / {len(df2)} rows -> frame of "
                    f"{len(res)} rows x {res.shape[1]} columns",
            "ms": dt * 1e3, "intervals_per_s": (len(df1) + len(df2)) / dt, "rows": int(len(res)),
            "frame_bytes": int(res.memory_usage(index=False).sum()),
        }
        del res
    else:
        out["cfg3_overlap_inner_frames"] = {"skipped": f"needs ~150 GB of free host memory, {room / 1e9:.0f} GB available"}
    return out


# ------------------------------------------------------------------------------------------ N > 1: range shards
def exchange_halo(q_local, rank, world, reach_of, dev):
    """Setup, outside the timed region (the partition of df1 into shards + halo rows is the host's job, SURVEY.md
    8e): every rank receives the queries of later ranks that a target of its own range can still reach --
    queries of its upper bound's chromosome with start below reach_of[r].  -> (grp, start, end, global row ids)
    of this rank's halo rows."""
    import torch.distributed as dist

    from bioframe_b200 import synth

    g, s, e = q_local
    n = len(s)
    send = {}
    for r in range(rank):  # earlier shards whose reach extends into this rank's rows
        g_hi, s_hi = synth.range_bounds(r, world)[2:]
        if g_hi >= N_GROUPS:
            continue
        m = np.flatnonzero((g == g_hi) & (s < reach_of[r]))
        if len(m):
            send[r] = (g[m], s[m], e[m], m.astype(np.int64) + rank * n)
    gathered = [None] * world
    dist.all_gather_object(gathered, send)
    parts = [gathered[r][rank] for r in range(world) if rank in gathered[r]]
    if not parts:
        z = np.empty(0, np.int64)
        return np.empty(0, np.int32), z, z.copy(), z.copy()
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(4))


def verify_shard(engine, dq, n_own, own_range, halo_ids, row_offset, dt, how, got):
    """The rows this rank produced against the single-GPU engine: bf_overlap_count/emit (its own sort, its own
    key layout) on the same local queries (own + halo) and all targets, with the rows another shard owns dropped
    -- an A-block pair (target starts at or after the query, arrops.py:338-339) belongs to the query's shard,
    a B-block pair (query starts after the target, arrops.py:341-342) to the target's, an unpaired row to the
    query's.  Dropping rows keeps the order, so the two must be equal row for row."""
    import torch

    ev1, ev2, _ = engine.overlap_intidxs(*dq, *dt, N_GROUPS, how=how)
    g1, s1 = dq[0], dq[1]
    g2, s2 = dt[0], dt[1]
    g_lo, s_lo, g_hi, s_hi = own_range

    def owned(g, s):
        return ((g > g_lo) | ((g == g_lo) & (s >= s_lo))) & ((g < g_hi) | ((g == g_hi) & (s < s_hi)))

    a, b = ev1.clamp(min=0), ev2.clamp(min=0)
    paired = ev2 >= 0
    is_a = paired & (s2[b] >= s1[a])
    own1 = owned(g1[a], s1[a])
    own2 = owned(g2[b], s2[b])
    keep = torch.where(is_a, own1, torch.where(paired, own2, own1))
    ev1, ev2 = ev1[keep], ev2[keep]
    want1 = torch.where(ev1 < n_own, ev1 + row_offset, halo_ids[(ev1 - n_own).clamp(min=0)] if halo_ids.numel() else ev1)
    return bool(torch.equal(want1, got[0]) and torch.equal(ev2, got[1]))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload not in ("overlap", "cfg4"):
        run_secondary(args)
        return

    import torch
    import torch.distributed as dist

    from bioframe_b200 import _lib, engine, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    cfg4 = args.workload == "cfg4"
    ranged = world > 1 or cfg4 or args.force_ranged
    chunks = max(args.chunks, 1) if cfg4 else 1

    # ---- synthetic workload, host side (pinned: the e2e leg copies from here) ----
    if cfg4:
        n1_total = args.n1 if args.n1 != 100_000_000 else 1_000_000_000
        n2 = args.n2 if args.n2 != 10_000_000 else 100_000_000
        n1 = n1_total // world
        q = synth.make_intervals_in_range(n1, 1 + 1000 * rank, "geom1k", rank, world, sort=True)
        t = synth.make_intervals(n2, 2, "geom1k") if rank == 0 else None
        desc = (f"_overlap_intidxs(how={args.how!r}) cfg4: {n1_total} hg38 queries (geom 1kb) pre-sorted by (chrom, start), "
                f"{n1} per GPU in {chunks} chunks, x {n2} targets (geom 1kb), sorted once on rank 0 and broadcast")
    elif ranged:
        n1 = args.n1 if args.scaling == "weak" else args.n1 // world
        n2 = args.n2
        q = synth.make_intervals_in_range(n1, 1 + 1000 * rank, "geom2k", rank, world)
        t = targets(n2)  # every rank draws them (the halo reach is computed from them); only rank 0's copy is used
        desc = workload_desc(args.how, n1, n2) + f"; queries of rank r inside the r-th of {world} (chrom, start) ranges"
    else:
        n1, n2 = args.n1, args.n2
        q, t = workload(n1, n2, rank)
        desc = workload_desc(args.how, n1, n2)
    row_offset = rank * n1
    own_range = synth.range_bounds(rank, world) if ranged else None

    halo = (np.empty(0, np.int32), np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.int64))
    lmax = 0
    if cfg4:  # the longest interval of either set bounds how far a target reaches beyond a range bound
        w, _ = engine.measure_extent(*[torch.from_numpy(a).to(dev) for a in (t if rank == 0 else q)], N_GROUPS)
        lmax = int(sharded.merge_extent_words(w, device=dev)[3])
        torch.cuda.empty_cache()
    if world > 1:
        if cfg4:  # the other ranks never hold the targets: reach = bound + longest interval (a superset)
            reach_of = [synth.range_bounds(r, world)[3] + lmax for r in range(world)]
        else:
            reach_of = [sharded.halo_reach(*t, synth.range_bounds(r, world)[2:], N_GROUPS) for r in range(world)]
        halo = exchange_halo(q, rank, world, reach_of, dev)
    n_halo = len(halo[1])
    q_all = tuple(np.concatenate([a, h]) for a, h in zip(q, halo[:3]))
    host_q = [torch.from_numpy(a).pin_memory() for a in q_all]
    have_targets = t is not None and (rank == 0 or not ranged)
    host_t = [torch.from_numpy(a).pin_memory() for a in t] if have_targets else None
    del q, q_all
    dq = [h.to(dev) for h in host_q]
    dt = [h.to(dev) for h in host_t] if have_targets else [None, None, None]
    halo_ids = torch.from_numpy(halo[3]).to(dev)
    h2d_bytes = sum(h.numel() * h.element_size() for h in host_q) + (sum(h.numel() * h.element_size() for h in host_t) if have_targets else 0)

    # cfg4: the rank's sorted queries in `chunks` contiguous sub-ranges (virtual shards); the halo of a chunk is
    # the head of the next one (local rows), the halo of the last chunk the rank's halo rows
    chunk_plan = None
    if cfg4 and chunks > 1:
        gq, sq = dq[0][:n1].cpu().numpy(), dq[1][:n1].cpu().numpy()
        chunk_plan = []
        cuts = [0] + [sharded._first_row_at_or_after(gq, sq, int(gq[n1 * c // chunks]), int(sq[n1 * c // chunks])) for c in range(1, chunks)] + [n1]
        for c in range(chunks):
            lo, hi = cuts[c], cuts[c + 1]
            g_lo, s_lo = (own_range[0], own_range[1]) if c == 0 else (int(gq[lo]), int(sq[lo]))
            g_hi, s_hi = (own_range[2], own_range[3]) if c == chunks - 1 else (int(gq[hi]), int(sq[hi]))
            if c == chunks - 1:
                h_hi = n1 + n_halo
            else:
                h_hi = sharded._first_row_at_or_after(gq, sq, g_hi, s_hi + lmax)
                h_hi = min(max(h_hi, hi), n1)
            chunk_plan.append(dict(lo=lo, hi=hi, h_hi=h_hi, own=(g_lo, s_lo, g_hi, s_hi)))
        del gq, sq

    last = {"pairs": 0, "rows": 0, "counts": None}

    def step(keep=False):
        if not ranged:
            a, b, npairs = engine.overlap_intidxs(dq[0], dq[1], dq[2], dt[0], dt[1], dt[2], N_GROUPS, how=args.how)
            last["pairs"], last["rows"] = npairs, int(a.numel())
            return a, b
        if chunk_plan is None:
            a, b, npairs, counts = sharded.overlap_range_sharded(
                dq[0], dq[1], dq[2], n1, own_range, dt[0], dt[1], dt[2], N_GROUPS, how=args.how, row_offset=row_offset,
                halo_ids=halo_ids, n2=n2, device=dev)
            last["pairs"], last["rows"], last["counts"] = npairs, int(a.numel()), counts
            return a, b
        # cfg4: targets prepared once per step, then one count/emit per chunk; the chunk results are released as
        # they are produced (130 GB of pairs per rank would not fit next to each other)
        words, pres = engine.measure_extent(dq[0], dq[1], dq[2], N_GROUPS)
        if rank == 0:
            w2, pres2 = engine.measure_extent(dt[0], dt[1], dt[2], N_GROUPS)
            words = np.maximum(words, w2)
        words = sharded.merge_extent_words(words, device=dev)
        side, main_s = torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
        side.wait_stream(main_s)
        handles = []
        with torch.cuda.stream(side):
            # persistent stores (engine.SetStores): a fresh multi-GB allocation per step that is used on two streams
            # goes back to the allocator through record_stream and costs a cudaMalloc every step
            set2 = (engine.prepare_set(words, dt[0], dt[1], dt[2], N_GROUPS, presorted=pres2, slot="cfg4.targets") if rank == 0
                    else engine.adopt_set(words, n2, N_GROUPS, slot="cfg4.targets"))
            if world > 1:
                handles = [dist.broadcast(set2.keys, src=0, async_op=True), dist.broadcast(set2.ids, src=0, async_op=True)]
            for h in handles:
                h.wait()  # the side stream waits for the sorted targets; the main stream sorts the first chunk meanwhile
            if args.how == "left":
                engine.set_runmax(set2)  # running max of the targets' ends: once per step, not once per chunk
        pairs = rows = 0
        for c, ch in enumerate(chunk_plan):
            sl = slice(ch["lo"], ch["h_hi"])
            set1 = engine.prepare_set(words, dq[0][sl], dq[1][sl], dq[2][sl], N_GROUPS, presorted=pres, slot="cfg4.queries")
            if c == 0:
                main_s.wait_stream(side)
            n_own = ch["hi"] - ch["lo"]
            local_halo = ch["h_hi"] - ch["hi"]
            hid = None
            if local_halo:
                hid = halo_ids if c == len(chunk_plan) - 1 else torch.arange(row_offset + ch["hi"], row_offset + ch["h_hi"], device=dev)
            a, b, npairs = engine.overlap_prepared(set1, set2, how=args.how, own_range=ch["own"], row_offset1=row_offset + ch["lo"],
                                                   n_own1=n_own, halo_ids=hid)
            pairs += npairs
            rows += int(a.numel())
            if not (keep and c == len(chunk_plan) - 1):
                del a, b
            del set1
        last["pairs"], last["rows"] = pairs, rows
        return (a, b) if keep else (None, None)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        ev1, ev2 = step()
        del ev1, ev2

    sampler = ClockSampler(local)
    sampler.start()
    _lib.prof_reset()
    _lib.prof_enable(True)
    fence()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        ev1, ev2 = step()
        del ev1, ev2
    t1.record()
    fence()
    sampler.stop_flag = True
    ms_total = t0.elapsed_time(t1)
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    n_pairs, n_rows = int(last["pairs"]), int(last["rows"])
    total_pairs, total_rows = n_pairs, n_rows
    if world > 1:
        tt = torch.tensor([n_pairs, n_rows], device=dev, dtype=torch.int64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        total_pairs, total_rows = int(tt[0].item()), int(tt[1].item())
    units = n1 * world + n2
    value = units / (ms_step * 1e-3)

    # ---- N > 1: every rank checks its rows against the single-GPU engine; the segment sizes of all ranks are
    #      exchanged once, after the last step (they place the segments in the whole result) ----
    sharded_verified = None
    if ranged and world > 1 and not args.no_verify and chunk_plan is None:
        got = step()
        counts_all = sharded.gather_group_counts(last["counts"], device=dev)
        _dst, _src, total = sharded.assemble_offsets(counts_all)
        full_t = [torch.from_numpy(a).to(dev) for a in t]  # regenerated from the seed on this rank, not the broadcast copy
        ok = verify_shard(engine, dq, n1, own_range, halo_ids, row_offset, full_t, args.how, got) and total == total_rows
        del got, full_t
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_verified = bool(flag.item())

    # cfg4: the same check on a virtual shard -- the first ~5e6 queries of this rank with their halo -- against the
    # single-GPU engine on all targets (sent to every rank once, after the timed region), and the row total of
    # the whole job against the independent count kernel
    if cfg4 and not args.no_verify:
        full_t = [torch.empty(n2, dtype=dt_, device=dev) for dt_ in (torch.int32, torch.int64, torch.int64)]
        if rank == 0:
            for dst_, src_ in zip(full_t, dt):
                dst_.copy_(src_)
        if world > 1:
            for x in full_t:
                dist.broadcast(x, src=0)
        gq, sq = dq[0][:n1].cpu().numpy(), dq[1][:n1].cpu().numpy()
        nv = min(5_000_000, n1 - 1)
        g_hi, s_hi = int(gq[nv]), int(sq[nv])
        nv = sharded._first_row_at_or_after(gq, sq, g_hi, s_hi)
        hv = min(max(sharded._first_row_at_or_after(gq, sq, g_hi, s_hi + lmax), nv), n1)
        own_v = (own_range[0], own_range[1], g_hi, s_hi)
        qv = [x[:hv] for x in dq]
        hid = torch.arange(row_offset + nv, row_offset + hv, device=dev)
        got = sharded.overlap_range_sharded(*qv, nv, own_v, *full_t, N_GROUPS, how=args.how, row_offset=row_offset,
                                            halo_ids=hid, n2=n2, device=dev, collective=False)[:2]
        ok = verify_shard(engine, qv, nv, own_v, hid, row_offset, full_t, args.how, got)
        del got
        cnt = engine.count_overlaps(dq[0][:n1], dq[1][:n1], dq[2][:n1], *full_t, N_GROUPS)
        want_rows = int(cnt.sum().item()) + (int((cnt == 0).sum().item()) if args.how == "left" else 0)
        del cnt, full_t
        tot = torch.tensor([want_rows, 1 if ok else 0], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        sharded_verified = bool(int(tot[1].item()) == world and int(tot[0].item()) == total_rows)
        engine.SetStores.release()
        torch.cuda.empty_cache()

    # ---- end to end: pinned host buffers in, host arrays out ----
    e2e = None
    if not args.no_e2e and not cfg4:
        # The result crosses PCIe as 4 bytes per row number and is widened into ordinary int64 host arrays by host
        # threads while the next chunk is in flight (bf_download_rows): half the bus traffic for the same int64
        # result.  With many ranks on one host the widening threads are scaled down to the cores each rank has;
        # below two threads the int64 columns are copied as they are (ring of two pinned 2-GiB buffers).
        threads = max(1, min(8, cores_available() // world - 1))
        narrow = threads >= 2
        direct_rows = [0]
        if narrow:
            # Result buffers reused by every step.  Page-locked when the host grants it: bf_download_rows then may
            # send the tail of a result straight into them on hosts whose threads are slower than the bus.
            try:
                res1, res2 = engine.pinned_rows(n_rows), engine.pinned_rows(n_rows)
                res_kind = "page-locked"
            except RuntimeError:
                res1, res2 = np.empty(n_rows, dtype=np.int64), np.empty(n_rows, dtype=np.int64)
                res_kind = "pageable"
            res1[:] = 0  # touched once here
            res2[:] = 0
        else:
            chunk = 1 << 28
            bufs = [torch.empty(chunk, dtype=torch.int64).pin_memory() for _ in range(2)]

        def e2e_step():
            for i, h in enumerate(host_q):
                dq[i].copy_(h, non_blocking=True)
            if have_targets:
                for i, h in enumerate(host_t):
                    dt[i].copy_(h, non_blocking=True)
            a, b = step()
            if narrow:
                engine.download_rows(a, out=res1, n_threads=threads)
                direct_rows[0] = engine.download_direct_rows()
                engine.download_rows(b, out=res2, n_threads=threads)
                direct_rows[0] += engine.download_direct_rows()
            else:
                k = 0
                for col in (a, b):
                    for lo in range(0, col.numel(), chunk):
                        part = col[lo:lo + chunk]
                        bufs[k & 1][: part.numel()].copy_(part, non_blocking=True)
                        k += 1
            torch.cuda.synchronize(dev)
            return a, b

        a, b = e2e_step()
        del a, b
        fence()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            a, b = e2e_step()
            del a, b
        fence()
        e2e_s = (time.perf_counter() - w0) / args.e2e_steps
        if world > 1:
            tt = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        e2e = {
            "value": units / e2e_s,
            "unit": "intervals/s",
            "ms_per_step": e2e_s * 1e3,
            "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": (8 * n_rows + 4 * direct_rows[0]) if narrow else 16 * n_rows,
            "result_bytes_per_step": 16 * n_rows,
            "api": "engine.overlap_intidxs / sharded.overlap_range_sharded on host columns (pre-encoded group codes); the "
                   "DataFrame-level call is timed in frame_e2e",
            "d2h_mode": (f"row numbers cross PCIe as int32 and are widened into {res_kind} int64 host arrays by {threads} "
                         f"host threads (bf_download_rows); {direct_rows[0]} of the {2 * n_rows} values of the last step went "
                         "as plain int64 copies instead (second lane, taken when the host threads are slower than the bus)"
                         ) if narrow else "int64 columns through a ring of two 2-GiB pinned buffers",
        }
        if narrow:
            a, b = step()  # the widened host arrays hold exactly what the device produced
            e2e["verified"] = bool(torch.equal(torch.from_numpy(res1), a.cpu()) and torch.equal(torch.from_numpy(res2), b.cpu()))
            del res1, res2, a, b
        else:
            del bufs

    if rank == 0:
        peak, peak_src = hbm_peak()
        sp = prof.get("sort_pass", {"ms": 0.0, "timed": 0, "bytes": 0})
        per_launch_ms = sp["ms"] / max(sp["timed"], 1)
        per_launch_bytes = sp["bytes"] / max(sp["timed"], 1)
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        launches = sum(v["launches"] for v in prof.values())
        # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of a 1e8-row launch,
        # scaled to this run's mix of launches (passes over n1 and over n2 rows, first pass reads no row id)
        traffic, traffic_src = None, None
        for name in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tj = json.load(f)
                npass = max(sp["timed"] // (2 * args.steps), 1)  # passes per set
                per_row_first, per_row_other = tj["first_pass_bytes"] / tj["rows"], tj["other_pass_bytes"] / tj["rows"]
                total_bytes = (n1 + n2) * (per_row_first + (npass - 1) * per_row_other)
                traffic = total_bytes / (2 * npass)
                traffic_src = f"profiles/{name} (ncu dram bytes of a 1e8-row launch, scaled per row)"
                break
            except Exception:
                continue
        n_sorted = n1 + n_halo + (n2 if (rank == 0 or not ranged) else 0)
        a_alg = 256 * n_sorted + 20 * n_rows
        a_alg_p4 = 160 * n_sorted + 20 * n_rows
        a_io = 20 * n_sorted + 16 * n_rows
        sec = ms_step * 1e-3
        line = {
            "metric": "bf.overlap input intervals/s",
            "value": value,
            "unit": "intervals/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "int64",
            "data": "synthetic",
            "config": {
                "workload": desc,
                "n1_per_gpu": n1,
                "n2": n2,
                "halo_rows_rank0": n_halo,
                "rows_out_per_gpu": n_rows,
                "pairs_out_per_gpu": n_pairs,
                "pairs_per_s": total_pairs / sec,
                "rows_out_total": total_rows,
                "l2": "inputs (2.2 GB) and every intermediate exceed the 126 MB L2; no flush needed",
                "parallelism": (f"queries sharded by (chrom, start) range x{world}, targets sorted on rank 0, sorted keys + ids "
                                "NCCL-broadcast on a side stream; row order = the reference's") if ranged else "single GPU",
            },
            "gpu_launches": launches // max(args.steps, 1) * args.steps,
            "clocks": sampler.summary(),
            "roofline": {
                "bound": "hbm",
                "kernel": "onesweep_pass_kernel<true> (radix pass, 8B key + 4B row id)",
                "achieved": achieved,
                "peak": peak,
                "peak_source": peak_src,
                "unit": "GB/s",
                "frac": achieved / peak if peak else None,
                "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "launches_timed": sp["timed"],
                "avg_launch_ms": per_launch_ms,
                "whole_step": {
                    "note": "A_alg is SURVEY.md 8d's stage model with p=8 radix passes; the code runs 4 (32-bit linear genome "
                            "axis), so A_alg_p4 = 160 N + 20 R is the model of what actually runs; A_io is the I/O floor",
                    "A_alg_GB": a_alg / 1e9,
                    "A_alg_p4_GB": a_alg_p4 / 1e9,
                    "A_io_GB": a_io / 1e9,
                    "achieved_alg_GBps": a_alg / sec / 1e9,
                    "frac_alg": a_alg / sec / 1e9 / peak,
                    "frac_alg_p4": a_alg_p4 / sec / 1e9 / peak,
                    "frac_io": a_io / sec / 1e9 / peak,
                },
                "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            },
            "e2e": e2e,
        }
        if sharded_verified is not None:
            line["sharded_verified"] = sharded_verified
        print_line = True

        def over_budget():
            return time.perf_counter() - _T_START > args.time_budget

        if world == 1 and not cfg4 and not args.no_cpu:
            leg = cpu_leg(args, 2, 1)
            line["cpu_baseline"] = {"value": leg["value"], "unit": "intervals/s", "cores": 1, "kind": leg["kind"],
                                    "sample": leg["desc"], "ms_per_step": leg["ms"]}
            # the GPU on the very same sample, row for row against the reference's result
            ds = [torch.from_numpy(a).to(dev) for a in (*leg["q"], *leg["t"])]
            g1_, g2_, _ = engine.overlap_intidxs(*ds, N_GROUPS, how=args.how)
            line["parity_checked"] = bool(np.array_equal(g1_.cpu().numpy(), leg["ev1"]) and np.array_equal(g2_.cpu().numpy(), leg["ev2"]))
            line["parity_sample"] = f"{leg['rows']} rows of the chr21+chr22 sample, array-equal with the {leg['kind']}, in its row order"
            del ds, g1_, g2_, leg
            # the same reference spread over every host core (what `--impl reference` times)
            if over_budget():
                line["cpu_baseline"]["all_cores"] = {"skipped": f"time budget ({args.time_budget:.0f} s)"}
            else:
                val_all, desc_all, ms_all, _, cores_all, kind_all = reference_leg_all_cores(args, 1, 1)
                line["cpu_baseline"]["all_cores"] = {"value": val_all, "cores": cores_all, "kind": kind_all, "sample": desc_all,
                                                     "ms_per_step": ms_all}
        # release the benchmark's tensors before the other legs
        del dq, dt, host_q, host_t
        engine.Workspace.release()
        torch.cuda.empty_cache()
        if world == 1 and not cfg4 and not args.no_frame:
            try:
                line["frame_e2e"] = {"skipped": f"time budget ({args.time_budget:.0f} s)"} if over_budget() else frame_leg(args, over_budget)
            except Exception as exc:  # a reported leg, never the reason the headline is lost
                line["frame_e2e"] = {"error": f"{type(exc).__name__}: {exc}"}
            engine.Workspace.release()
            torch.cuda.empty_cache()
        if world == 1 and not cfg4 and not args.no_secondary:
            sec_out = {}
            for wl in ("cfg0", "closest", "cluster", "coverage", "ordered"):
                if over_budget():
                    sec_out[wl] = {"skipped": f"time budget ({args.time_budget:.0f} s)"}
                    continue
                try:
                    s_line = secondary_line(args, wl, 3, 3, with_cpu=not args.no_cpu)
                    sec_out[wl] = {"metric": s_line["metric"], "workload": s_line["config"]["workload"], "ms_per_step": s_line["ms_per_step"],
                                   "value": s_line["value"], "rows_out": s_line["config"]["rows_out"],
                                   "roofline_frac": s_line["roofline"]["frac"], "gpu_launches": s_line["gpu_launches"],
                                   "cpu_baseline": s_line.get("cpu_baseline")}
                except Exception as exc:
                    sec_out[wl] = {"error": f"{type(exc).__name__}: {exc}"}
                engine.Workspace.release()
                torch.cuda.empty_cache()
            line["secondary"] = sec_out
        if print_line:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
