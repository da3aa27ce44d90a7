#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`):
per kernel launches, total time, share and -- when present -- DRAM traffic.

  python tools/launch_summary.py LAUNCHES.csv [STEPS] [--json OUT.json]
"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
rows = list(csv.reader(open(path, errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mu, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0}
agg = collections.defaultdict(lambda: {"n": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("b200ssl::", "")[:64]
    try:
        v = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)
    except ValueError:
        continue
    a = agg[name]
    a["n"].add(r[idc])
    if r[mn].startswith("gpu__time_duration"):
        a["ns"] += v
    elif r[mn].startswith("dram__bytes_read"):
        a["rd"] += v
    elif r[mn].startswith("dram__bytes_write"):
        a["wr"] += v
tot = sum(a["ns"] for a in agg.values())
has_dram = any(a["rd"] or a["wr"] for a in agg.values())
print(f"{'kernel':64s} {'launches/step':>13s} {'ms/step':>9s} {'share':>6s}" + (f" {'DRAM MB/step':>13s} {'GB/s':>7s}" if has_dram else ""))
out = []
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[:48]:
    line = f"{k:64s} {len(a['n'])/steps:13.1f} {a['ns']/1e6/steps:9.3f} {100*a['ns']/tot:5.1f}%"
    if has_dram:
        line += f" {(a['rd']+a['wr'])/1e6/steps:13.1f} {(a['rd']+a['wr'])/max(a['ns'],1):7.0f}"
    print(line)
    out.append({"kernel": k, "launches_per_step": len(a["n"]) / steps, "ms_per_step": a["ns"] / 1e6 / steps,
                "share": a["ns"] / tot, "dram_read_mb_per_step": a["rd"] / 1e6 / steps,
                "dram_write_mb_per_step": a["wr"] / 1e6 / steps})
print(f"{'TOTAL':64s} {sum(len(a['n']) for a in agg.values())/steps:13.1f} {tot/1e6/steps:9.3f}")
if "--json" in sys.argv:
    json.dump({"source": path, "steps": steps, "kernels": out, "total_ms_per_step": tot / 1e6 / steps},
              open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
