#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) by kernel:
launches, total time, share.  Usage: python tools/launch_shares.py X.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    k = r[ki].split("(")[0][:70]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(t for _, t in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {t:.1f} | {t / tot:.1%} |")
