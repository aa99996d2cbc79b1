import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early
#!/usr/bin/env python
"""Aggregate an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum[,more] --csv`) by kernel.
Only rows of the metric gpu__time_duration.sum are summed (a csv captured with several metrics has one row
per metric per launch), converted from the unit the row states."""
import collections
import csv
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    col = {c: i for i, c in enumerate(rows[0])}
    agg = collections.OrderedDict()
    tot, n_launch = 0.0, 0
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[col["Kernel Name"]][:86]
        t = float(r[col["Metric Value"]].replace(",", "")) * UNIT[r[col["Metric Unit"]]]
        a = agg.setdefault(k, [0, 0.0, r[col["Grid Size"]], r[col["Block Size"]]])
        a[0] += 1
        a[1] += t
        tot += t
        n_launch += 1
    print(f"== {path}: {n_launch} launches, {tot:.1f} us")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"{n:4d} {t:10.1f} us  avg {t / n:8.2f}  {100 * t / tot:5.1f}%  {g:>14s} {b:>12s} {k}")
