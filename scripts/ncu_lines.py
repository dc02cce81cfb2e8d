#!/usr/bin/env python
"""Per source line: warp instructions, stall samples and shared-memory wavefronts, from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > X.csv   (all files of the kernel)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hdr = None
cur = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            inst = int(d["Instructions Executed"])
            smp = int(d["# Samples"])
        except ValueError:
            continue
        key = (cur.split("/")[-1], int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += inst
        a[1] += smp
        try:
            a[2] += int(d["L1 Wavefronts Shared"])
        except (ValueError, KeyError):
            pass
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp wf=%11d %s:%d %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2], k[0], k[1], k[2]))
