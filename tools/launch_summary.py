#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):  python tools/launch_summary.py <csv> [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + skip:]:
    if len(r) <= vi:
        continue
    k = r[ki].split("(")[0][-60:]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%5d total %11.1f us  avg %10.1f us  %5.1f%%" % (k, n, t, t / n, 100 * t / tot))
