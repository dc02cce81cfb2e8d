#!/usr/bin/env python
"""Print the interesting numbers of bench JSON lines: python scripts/show.py gpurun_out/x.json ..."""
import json
import sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, "ERR", e)
        continue
    k = d.get("roofline", {}).get("kernels", {})
    e2e = d.get("e2e") or {}
    print("%-40s %7.2f G/s %8.2f ms/step e2e %s check %s" % (fn.split("/")[-1], d["value"] / 1e9, d["ms_per_step"],
          ("%.2f" % (e2e["value"] / 1e9)) if e2e.get("value") else None, (d.get("check") or {}).get("tables_checksum_equal_reference")))
    for a, b in k.items():
        print("      %-40s %8.3f ms/launch x %d  share %.2f" % (a, b.get("ms_per_launch") or (b["ms_total"] / max(1, b["launches"])), b["launches"], b.get("share_of_step", 0)))
