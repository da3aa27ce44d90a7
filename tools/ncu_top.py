#!/usr/bin/env python
"""Developer helper: top SASS instructions by warp-stall samples for one kernel launch of an .ncu-rep.

  python tools/ncu_top.py REPORT.ncu-rep LAUNCH_INDEX [TOP_N] [CONTEXT]
"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
all_lines = out.splitlines()
starts = [i for i, l in enumerate(all_lines) if l.startswith('"Kernel Name"')] + [len(all_lines)]
lines = all_lines[starts[idx]:starts[idx + 1]]
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
h = rows[0]
ci = {n: i for i, n in enumerate(h)}
S, SRC, EX = ci["# Samples"], ci["Source"], ci["Instructions Executed"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[1:] if len(r) == len(h)]
tot = sum(int(r[S]) for r in body)
tot_ex = sum(int(r[EX]) for r in body)
print(f"total samples {tot}, warp instructions executed {tot_ex}")
agg = {}
for n, i in stall_cols:
    agg[n] = sum(int(r[i] or 0) for r in body)
print("stall totals:", ", ".join(f"{n[6:]}={v}" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda k: -int(body[k][S]))[:top_n]
for k in order:
    r = body[k]
    st = sorted(((int(r[i] or 0), n[6:]) for n, i in stall_cols), reverse=True)[:3]
    print(f"--- #{k} samples {r[S]} ({100*int(r[S])/max(tot,1):.1f}%) exec {r[EX]}  {', '.join(f'{n}={v}' for v, n in st if v)}")
    for j in range(max(0, k - ctx), min(len(body), k + ctx + 1)):
        print(f"   {'>>' if j == k else '  '} {body[j][SRC].strip()[:110]}   [{body[j][S]}]")

if "--mix" in sys.argv:
    import collections
    import re
    mix = collections.Counter()
    for r in body:
        op = r[SRC].strip().split()
        if not op:
            continue
        name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        mix[name.split(".")[0]] += int(r[EX])
    print("instruction mix (warp instructions executed):")
    for name, v in mix.most_common(30):
        print(f"   {name:12s} {v:12d}  {100*v/tot_ex:5.1f}%")
