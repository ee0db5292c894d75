#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    name = row["Kernel Name"]
    name = name.replace("void ", "").replace("<unnamed>::", "")
    agg[name[:90]][0] += 1
    agg[name[:90]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'ms/step':>9} {'launches/step':>13} {'share':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6 / steps:9.3f} {v[0] / steps:13.1f} {100 * v[1] / tot:5.1f}%  {k}")
print(f"{tot / 1e6 / steps:9.3f} total ms/step (serialised, cold-cache ncu timings)")
