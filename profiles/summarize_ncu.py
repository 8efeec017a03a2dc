#!/usr/bin/env python
"""Text summary of an .ncu-rep (ncu --set full): headline metrics per captured launch + the hottest SASS
lines by stall samples.  usage: summarize_ncu.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## launch id {d.get('ID')}: {d.get('Kernel Name', '')[:110]}")
        for w in WANT:
            if w in d:
                print(f"  {w:82s} {d[w]:>16s} {units[hdr.index(w)]}")
        try:
            rd = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
            t = float(d["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
            print(f"  => DRAM traffic {(rd + wr) / 1e9:.3f} GB in {t * 1e3:.3f} ms = {(rd + wr) / t / 1e9:.0f} GB/s (under ncu, cold cache)")
        except Exception:
            pass
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    h, out = None, []
    kernels = 0
    for r in src:
        if r and r[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break
        if r and r[0] == "Address":
            h = r
            continue
        if h and len(r) == len(h):
            d = dict(zip(h, r))
            try:
                d["_n"] = int(d["# Samples"])
            except ValueError:
                continue
            out.append(d)
    total = sum(d["_n"] for d in out) or 1
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    print(f"\n## hottest SASS lines of the first launch ({total} stall samples, {len(out)} instructions)")
    for d in sorted(out, key=lambda d: -d["_n"])[:16]:
        why = sorted(((int(d[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"  {100 * d['_n'] / total:5.1f}%  {d['Source'][:84]:84s} {why}")


if __name__ == "__main__":
    main()
