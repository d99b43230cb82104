#!/usr/bin/env python
"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last STEP's kernels in launch order
with their durations, and the sum.  A step is recognised by the first kernel name repeating.
usage: tools/launch_table.py profiles/r2_launches_refscene_25k_warm_final.csv [first-kernel-substring]"""
import csv
import sys

path = sys.argv[1]
first = sys.argv[2] if len(sys.argv) > 2 else "bh_bbox_kernel"
rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
starts = [i for i, r in enumerate(body) if first in r[ki]]
if len(starts) < 2:
    sys.exit(f"fewer than two launches of a kernel matching {first!r}")
step = body[starts[-2]:starts[-1]]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
total = 0.0
for r in step:
    us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)
    total += us
    print(f"{us:9.2f} us  {r[ki].split('(')[0][:90]}")
print(f"{total:9.2f} us  sum of {len(step)} launches")
