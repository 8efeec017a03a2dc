#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/ncu_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = []
hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        if out:
            break  # first kernel only
        print("KERNEL", r[1][:120])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            d["_n"] = int(d["# Samples"])
        except ValueError:
            continue
        out.append(d)
total = sum(d["_n"] for d in out) or 1
print("total samples", total, "instructions", len(out))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for d in sorted(out, key=lambda d: -d["_n"])[:top]:
    why = sorted(((int(d[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{d['_n']:6d} {100*d['_n']/total:5.1f}%  {d['Source'][:90]:90s} {why}")
