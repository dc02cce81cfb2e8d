#!/usr/bin/env python
"""Executed warp instructions by SASS opcode, from  ncu -i X.ncu-rep --page source --csv --print-source sass > X.csv"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
tot = 0
byop = {}
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        toks = d.get("Source", "").split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        op = op.split(".")[0] if len(sys.argv) > 2 and sys.argv[2] == "short" else op
        byop[op] = byop.get(op, 0) + n
        tot += n
print("total", tot)
for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%6.2f%% %s" % (100 * v / tot, k))
