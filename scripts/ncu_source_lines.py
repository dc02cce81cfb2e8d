#!/usr/bin/env python
"""Per source line: warp instructions executed + stall samples, from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > X.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = []
for r in rows:
    if len(r) >= 8 and r[0].isdigit():
        try:
            lines.append((int(r[7]), int(r[6]), int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(x[0] for x in lines) or 1
tots = sum(x[1] for x in lines) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
for n, smp, ln, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  L%-4d %s" % (100.0 * n / tot, 100.0 * smp / tots, ln, src[:120]))
