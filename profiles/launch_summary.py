#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST
bench step (the list holds warm-up steps too).  usage: launch_summary.py launches.csv [n_steps_in_list]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
per = len(data) // steps
last = data[(steps - 1) * per:]
agg = collections.OrderedDict()
for d in last:
    k = d["Kernel Name"].split("(")[0][:48] + " grid=" + d["Grid Size"]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += float(d["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches in list, {len(last)} in the last step, {tot:.3f} ms (cold-cache, serialised)")
for k, v in agg.items():
    print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:<3d} {k}")
