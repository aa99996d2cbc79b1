#!/usr/bin/env python
"""Aggregate an ncu launch list (csv from --metrics gpu__time_duration.sum) by kernel."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    col = {c: i for i, c in enumerate(rows[0])}
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        k, t = r[col["Kernel Name"]][:78], float(r[col["Metric Value"]]) / 1e3
        a = agg.setdefault(k, [0, 0.0, r[col["Grid Size"]], r[col["Block Size"]]])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"== {path}: {len(rows) - 1} launches, {tot:.1f} us")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{n:4d} {t:10.1f} us  avg {t / n:8.2f}  {g:>14s} {b:>12s} {k}")
