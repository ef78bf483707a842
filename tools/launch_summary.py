#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py list.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0][:64]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    acc.setdefault(n, [0, 0.0])
    acc[n][0] += 1
    acc[n][1] += v
tot = sum(v for _, v in acc.values())
for n, (c, v) in sorted(acc.items(), key=lambda x: -x[1][1]):
    print(f"{n:66s} {c / steps:6.1f} launches {v / steps:10.1f} us  {100 * v / tot:5.1f} %")
print(f"total {tot / steps:.1f} us per step")
