#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv "command line" > profiles/xx.md"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[h.index("Metric Unit")]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += v
tot = sum(v[1] for v in agg.values())
print("# ncu launch list (cold-cache, serialised: compare SHARES, not absolutes)\n")
print("command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv %s`\n" % sys.argv[2])
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k[:90], n, t, 100 * t / tot))
