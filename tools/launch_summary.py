#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.

  python tools/launch_summary.py LAUNCHES.csv [STEPS]      (STEPS: number of steps the list covers; per-step figures)
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path, errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    name = name.replace("void ", "").replace("b200ssl::", "")[:64]
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':64s} {'launches/step':>13s} {'ms/step':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{k:64s} {v[0]/steps:13.1f} {v[1]/1e6/steps:9.3f} {100*v[1]/tot:5.1f}%")
print(f"{'TOTAL':64s} {sum(v[0] for v in agg.values())/steps:13.1f} {tot/1e6/steps:9.3f}")
