#!/usr/bin/env python
"""Print the headline keys of one bench.py JSON line (tools/gpu.sh bench)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value", d.get("value"), d.get("unit"), "ms_step", d.get("ms_per_step"), "launches", d.get("gpu_launches"))
for k in ("roofline", "roofline_other"):
    r = d.get(k)
    if r:
        print(k, r.get("kernel"), "ms", r.get("ms_per_launch"), "bound", r.get("bound"), "achieved", r.get("achieved"),
              r.get("unit"), "frac", r.get("frac"), "share", r.get("share_of_step"))
for k in ("e2e", "e2e_model", "cpu_baseline", "clocks"):
    if k in d:
        print(k, json.dumps(d[k])[:400])
rows = d.get("rows", [])
for r in (rows if isinstance(rows, list) else [rows]):
    print(json.dumps(r)[:600])
