#!/usr/bin/env python
"""bench.py -- the north-star benchmark: `bf.overlap` index path at 1e8 x 1e7.

One "step" = one pass of `_overlap_intidxs(how='left')` over
the cfg3 workload of BASELINE.json: 1e8 synthetic hg38 queries (Geometric mean
2 kb) against 1e7 peak-like targets, ~9 overlaps per query, ~9e8 output rows.

  value       (N1*n_gpus + N2) / t, inputs resident in HBM, CUDA-event timed
  e2e         same call with HOST (pinned) buffers: H2D of the three columns of
              both sets and D2H of both int64 result columns inside the timed region
  roofline    dominant kernel = one onesweep radix pass over the query keys:
              24 B/row algorithmic (8 B key + 4 B row id, read and written)
  cpu_baseline the numpy oracle port of the reference algorithm on a bounded,
              same-density sample (chr21+chr22 share of both sets), 1 core

`--impl reference` times only that CPU port (there is no other reference
implementation of this path: the reference is pure numpy/pandas).

N > 1 (torchrun): queries are sharded (each rank owns 1e8 of them, weak
scaling), the target columns are broadcast from rank 0 over NCCL inside every
step, no other collective.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="overlap", choices=["overlap", "closest", "cluster", "coverage", "count", "subtract", "ingest", "ordered"],
                    help="overlap = the north-star line (cfg3); the others are the secondary BASELINE.json configs "
                         "(closest: cfg1 1e7 x 1e6 k=1; cluster: cfg2 1e8 heavy-tailed; coverage: cfg3 sets)")
    ap.add_argument("--k", type=int, default=1)
    return ap.parse_args()


def workload(n1, n2, rank, chrom_ids=None):
    from bioframe_b200 import synth

    frac = synth.subset_fraction(chrom_ids) if chrom_ids is not None else 1.0
    m1, m2 = max(int(n1 * frac), 1), max(int(n2 * frac), 1)
    q = synth.make_intervals(m1, 1 + 1000 * rank, "geom2k", chrom_ids=chrom_ids)
    hot = max(int(200_000 * frac * (n2 / 1e7)), 1)
    t = synth.make_peak_targets(m2, 2, n_hotspots=hot, chrom_ids=chrom_ids)
    return q, t


def workload_desc(how, n1, n2):
    return (f"_overlap_intidxs(how={how!r}) cfg3: {n1} hg38 queries (geom 2kb) per GPU x {n2} peak-like targets, "
            "unsorted")


def cpu_leg(args, steps, warmup):
    """Oracle port on the bounded sample; returns (intervals/s, description, ms/step, rows)."""
    from oracle import ops_oracle as oo

    (g1, s1, e1), (g2, s2, e2) = workload(args.n1, args.n2, 0, chrom_ids=SAMPLE_CHROMS)
    times = []
    rows = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ev1, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, N_GROUPS, how=args.how)
        dt = time.perf_counter() - t0
        rows = int(ev1.shape[0])
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    n = len(s1) + len(s2)
    desc = (
        f"chr21+chr22 share of the workload at full density: {len(s1)} x {len(s2)} intervals -> "
        f"{rows} rows per step, numpy oracle port of ops._overlap_intidxs(how={args.how!r})"
    )
    return n / t, desc, t * 1e3, rows


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


def _reference_worker(conn, n1, n2, how, chrom):
    """One process of the reference arm: the per-chromosome body of the reference's loop (ops.py:301-336) on one
    chromosome of the workload at full density, run whenever the parent says go."""
    from oracle import ops_oracle as oo

    (g1, s1, e1), (g2, s2, e2) = workload(n1, n2, 0, chrom_ids=(chrom,))
    conn.send((len(s1), len(s2)))
    while conn.recv():
        ev1, _ = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, N_GROUPS, how=how)
        conn.send(int(ev1.shape[0]))


def reference_leg_all_cores(args, steps, warmup):
    """The reference's algorithm on every host core: its chromosome groups are independent, so one process per
    chromosome is how its user spreads it over a machine.  The `cores` smallest chromosomes of the workload at
    full density, one per process, all started together; a step ends when the slowest process is done.
    Returns (intervals/s, description, ms/step, rows, cores)."""
    import multiprocessing as mp

    from bioframe_b200 import synth

    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, N_GROUPS))
    chroms = [int(c) for c in np.argsort(synth.HG38_SIZES)[:cores]]
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
    procs = []
    for c in chroms:
        parent, child = ctx.Pipe()
        pr = ctx.Process(target=_reference_worker, args=(child, args.n1, args.n2, args.how, c), daemon=True)
        pr.start()
        procs.append((pr, parent))
    sizes = [conn.recv() for _, conn in procs]
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
    t = float(np.mean(times))
    n = sum(a + b for a, b in sizes)
    names = ",".join(synth.HG38_CHROMS[c] for c in chroms)
    desc = (
        f"{cores} processes, one chromosome each ({names}) of the workload at full density: "
        f"{sum(a for a, _ in sizes)} x {sum(b for _, b in sizes)} intervals -> {rows} rows per step, numpy oracle port "
        f"of the per-chromosome body of ops._overlap_intidxs(how={args.how!r})"
    )
    return n / t, desc, t * 1e3, rows, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, desc, ms, rows, cores = reference_leg_all_cores(args, args.steps, args.warmup)
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
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": workload_desc(args.how, args.n1, args.n2), "n1_per_gpu": args.n1, "n2": args.n2,
                   "sample": desc},
        "cpu_baseline": {"value": val, "unit": "intervals/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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


def run_secondary(args):
    """closest / cluster / coverage at their BASELINE.json sizes, single GPU: device-resident inputs,
    CUDA-event timing of the whole call, whole-call algorithmic bytes against the measured
    HBM peak, and the oracle port on a bounded sample beside it."""
    import torch

    from bioframe_b200 import _lib, engine, synth
    from oracle import ops_oracle as oo

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = args.workload
    cpu_sample = None
    if wl == "closest":
        n1, n2 = (args.n1 if args.n1 != 100_000_000 else 10_000_000), (args.n2 if args.n2 != 10_000_000 else 1_000_000)
        q = synth.make_intervals(n1, 1, "geom1k")
        t = synth.make_intervals(n2, 2, "geom1k")
        d = [torch.from_numpy(a).to(dev) for a in (*q, *t)]
        call = lambda: engine.closest_intidxs(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS, k=args.k)  # noqa: E731
        rows_of = lambda r: int(r[0].numel())  # noqa: E731
        a_alg = lambda R: 2 * 256 * n2 + 36 * n1 + 20 * R  # noqa: E731
        units, name = n1 + n2, f"bf.closest(k={args.k}) input intervals/s"
        desc = f"_closest_intidxs cfg1: {n1} hg38 queries x {n2} targets (geom 1kb), k={args.k}, unsorted"
        frac = synth.subset_fraction(SAMPLE_CHROMS)
        sq = synth.make_intervals(int(n1 * frac), 1, "geom1k", chrom_ids=SAMPLE_CHROMS)
        stt = synth.make_intervals(int(n2 * frac), 2, "geom1k", chrom_ids=SAMPLE_CHROMS)
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
        frac = synth.subset_fraction(SAMPLE_CHROMS)
        sq = synth.make_intervals(int(n1 * frac), 1, "pareto", chrom_ids=SAMPLE_CHROMS)
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
            a, b = oo.overlap_intidxs_coded(*sq, *stt, N_GROUPS, how="left")
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
        cpu = lambda: oo.coverage_coded(*sq, *stt, N_GROUPS)  # noqa: E731
        cpu_units = len(sq[1]) + len(stt[1])

    for _ in range(max(args.warmup, 3)):
        r = call()
    R = rows_of(r)
    del r
    _lib.prof_reset()
    _lib.prof_enable(True)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        r = call()
        del r
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1) / args.steps
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    peak = 6554.6
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    alg = a_alg(R)
    line = {
        "metric": name, "value": units / (ms * 1e-3), "unit": "intervals/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc, "rows_out": R, "l2": "inputs and intermediates exceed the 126 MB L2"},
        "gpu_launches": sum(v["launches"] for v in prof.values()),
        "roofline": {"bound": "hbm", "kernel": "whole call (A_alg of SURVEY.md 8d)", "achieved": alg / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}},
    }
    if not args.no_cpu:
        times = []
        for it in range(2):
            w0 = time.perf_counter()
            cpu()
            times.append(time.perf_counter() - w0)
        line["cpu_baseline"] = {"value": cpu_units / min(times), "unit": "intervals/s", "cores": 1,
                                "kind": "reference" if wl == "ingest" else "port",
                                "sample": cpu_sample or f"chr21+chr22 share of the workload at full density ({cpu_units} intervals), numpy oracle port",
                                "ms_per_step": min(times) * 1e3}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload != "overlap":
        run_secondary(args)
        return

    import torch
    import torch.distributed as dist

    from bioframe_b200 import _lib, engine, sharded

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

    # ---- synthetic workload, host side (pinned: the e2e leg copies from here) ----
    (g1, s1, e1), (g2, s2, e2) = workload(args.n1, args.n2, rank)
    host = [torch.from_numpy(a).pin_memory() for a in (g1, s1, e1, g2, s2, e2)]
    n1, n2 = len(s1), len(s2)
    del g1, s1, e1, g2, s2, e2  # the pinned copies are the host side from here on
    d = [h.to(dev) for h in host]
    if world > 1 and rank != 0:
        for t in d[3:]:
            t.zero_()  # targets arrive by broadcast
    h2d_bytes = sum(h.numel() * h.element_size() for h in host)

    row_offset = rank * n1  # weak scaling: rank r owns query rows [r*n1, (r+1)*n1) of the whole job

    def step():
        if world > 1:
            # targets broadcast from rank 0 over NCCL/NVLink inside the step, queries stay local,
            # one all-gather of the per-rank row counts (bioframe_b200/sharded.py)
            ev1, ev2, _off, _tot = sharded.overlap_sharded(d[0], d[1], d[2], row_offset, d[3], d[4], d[5], N_GROUPS,
                                                           how=args.how)
            return ev1, ev2, last_pairs[0]
        a, b, npairs = engine.overlap_intidxs(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS, how=args.how)
        last_pairs[0] = npairs
        return a, b, npairs

    last_pairs = [0]
    if world > 1:  # pair count of this rank's shard (for pairs/s), measured once outside the timed region
        _a, _b, last_pairs[0] = engine.overlap_intidxs(d[0], d[1], d[2], d[3], d[4], d[5], N_GROUPS, how=args.how)
        del _a, _b

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        ev1, ev2, n_pairs = step()
    n_rows = int(ev1.numel())
    del ev1, ev2

    sampler = ClockSampler(local)
    sampler.start()
    _lib.prof_reset()
    _lib.prof_enable(True)
    fence()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        ev1, ev2, n_pairs = step()
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
    total_pairs, total_rows = int(n_pairs), n_rows
    if world > 1:
        tt = torch.tensor([int(n_pairs), n_rows], device=dev, dtype=torch.int64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        total_pairs, total_rows = int(tt[0].item()), int(tt[1].item())
    units = n1 * world + n2
    value = units / (ms_step * 1e-3)

    # ---- end to end: pinned host buffers in, pinned host buffers out ----
    e2e = None
    if not args.no_e2e:
        # Host memory: two pinned int64 result columns are 2 x 7 GB per rank (the pinned allocator rounds up to
        # 8 GB each).  Up to 4 ranks they are kept whole; with more ranks on one box (196 GB of host RAM here)
        # the result is read back through a ring of two 2-GiB pinned buffers instead -- same bytes over PCIe.
        # With one or two ranks on the box the result crosses PCIe as 4 bytes per row number and is widened into
        # ordinary int64 host arrays by host threads while the next chunk is in flight (bf_download_rows): half
        # the bus traffic for the same int64 result (315 -> 198 ms at one GPU).  The widening doubles the host
        # memory traffic, though, and with more ranks sharing one host the memory controllers, not PCIe, are
        # the limit (440 vs 467 ms at two GPUs: already a tie): from four ranks on the int64 columns are
        # copied as they are into pinned host arrays.
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        threads = max(1, min(8, cores // world - 1))
        narrow = threads >= 4 and world <= 2
        ring = (not narrow) and world > 4
        if narrow:
            res1 = np.zeros(n_rows, dtype=np.int64)  # pageable, touched once here
            res2 = np.zeros(n_rows, dtype=np.int64)
        elif ring:
            chunk = 1 << 28
            bufs = [torch.empty(chunk, dtype=torch.int64).pin_memory() for _ in range(2)]
        else:
            out1 = torch.empty(n_rows, dtype=torch.int64).pin_memory()
            out2 = torch.empty(n_rows, dtype=torch.int64).pin_memory()

        def e2e_step():
            dd = [h.to(dev, non_blocking=True) for h in host]
            a, b, _ = engine.overlap_intidxs(dd[0], dd[1], dd[2], dd[3], dd[4], dd[5], N_GROUPS, how=args.how)
            if narrow:
                engine.download_rows(a, out=res1, n_threads=threads)
                engine.download_rows(b, out=res2, n_threads=threads)
            elif ring:
                k = 0
                for col in (a, b):
                    for lo in range(0, col.numel(), chunk):
                        part = col[lo:lo + chunk]
                        bufs[k & 1][: part.numel()].copy_(part, non_blocking=True)
                        k += 1
            else:
                out1[: a.numel()].copy_(a, non_blocking=True)
                out2[: b.numel()].copy_(b, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        fence()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
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
            "d2h_bytes_per_step": (8 if narrow else 16) * n_rows,
            "result_bytes_per_step": 16 * n_rows,
            "d2h_mode": (f"row numbers cross PCIe as int32 and are widened into int64 host arrays by {threads} host threads "
                         "(bf_download_rows)") if narrow
            else ("ring of two 2-GiB pinned buffers" if ring else "whole result into pinned host arrays"),
        }
        if narrow:
            # the widened host arrays hold exactly what the device produced
            a, b, _ = engine.overlap_intidxs(*[h.to(dev) for h in host], N_GROUPS, how=args.how)
            e2e["verified"] = bool(torch.equal(torch.from_numpy(res1), a.cpu()) and torch.equal(torch.from_numpy(res2), b.cpu()))
            del res1, res2, a, b
        elif ring:
            del bufs
        else:
            del out1, out2

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        sp = prof.get("sort_pass", {"ms": 0.0, "timed": 0, "bytes": 0})
        per_launch_ms = sp["ms"] / max(sp["timed"], 1)
        per_launch_bytes = sp["bytes"] / max(sp["timed"], 1)
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        launches = sum(v["launches"] for v in prof.values())
        # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of a 1e8-row launch,
        # scaled to this run's mix of launches (passes over n1 and over n2 rows, first pass reads no row id)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            npass = max(sp["timed"] // (2 * args.steps), 1)  # passes per set
            per_row_first, per_row_other = tj["first_pass_bytes"] / tj["rows"], tj["other_pass_bytes"] / tj["rows"]
            total_bytes = (n1 + n2) * (per_row_first + (npass - 1) * per_row_other)
            traffic = total_bytes / (2 * npass)
        except Exception:
            pass
        a_alg = 256 * (n1 + n2) + 20 * n_rows
        a_io = 20 * (n1 + n2) + 16 * n_rows
        line = {
            "metric": "bf.overlap input intervals/s",
            "value": value,
            "unit": "intervals/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int64",
            "data": "synthetic",
            "config": {
                "workload": workload_desc(args.how, n1, n2),
                "n1_per_gpu": n1,
                "n2": n2,
                "rows_out_per_gpu": n_rows,
                "pairs_out_per_gpu": int(n_pairs),
                "pairs_per_s": total_pairs / (ms_step * 1e-3),
                "rows_out_total": total_rows,
                "l2": "inputs (2.2 GB) and every intermediate exceed the 126 MB L2; no flush needed",
                "parallelism": f"query-sharded x{world}, targets NCCL-broadcast" if world > 1 else "single GPU",
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
                "traffic_source": "profiles/r1_traffic.json (ncu dram bytes of a 1e8-row launch, scaled per row)",
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "launches_timed": sp["timed"],
                "avg_launch_ms": per_launch_ms,
                "whole_step": {
                    "A_alg_GB": a_alg / 1e9,
                    "A_io_GB": a_io / 1e9,
                    "achieved_alg_GBps": a_alg / (ms_step * 1e-3) / 1e9,
                    "achieved_io_GBps": a_io / (ms_step * 1e-3) / 1e9,
                    "frac_alg": a_alg / (ms_step * 1e-3) / 1e9 / peak,
                    "frac_io": a_io / (ms_step * 1e-3) / 1e9 / peak,
                },
                "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            },
            "e2e": e2e,
        }
        if world == 1 and not args.no_cpu:
            val, desc, ms, _ = cpu_leg(args, 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": "intervals/s", "cores": 1, "kind": "port", "sample": desc, "ms_per_step": ms}
            # the same port spread over every host core, one chromosome per process (what `--impl reference` times)
            val_all, desc_all, ms_all, _, cores_all = reference_leg_all_cores(args, 2, 1)
            line["cpu_baseline"]["all_cores"] = {"value": val_all, "cores": cores_all, "sample": desc_all, "ms_per_step": ms_all}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
