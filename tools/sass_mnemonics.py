#!/usr/bin/env python
"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass):
UTCHMMA/UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG / UTMASTG / UTMAREDG
(TMA tensor load / store / reduce), UBLKCP (1-D bulk copy), SYNCS (mbarrier), plus HMMA (mma.sync: must be 0).

  python tools/sass_mnemonics.py [LIB.so] > profiles/r2_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gipmed-project-self-supervised-vit_b200", "libb200ssl.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA",
             "FFMA2", "MUFU.EX2", "RED.E"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[cur][mn] += 1
names = list(counts)
try:
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
except Exception:
    pass


def short(n):
    d = demangle.get(n, n)
    d = re.sub(r"\(.*", "", d).replace("void ", "").replace("b200ssl::", "")
    return d[:58]


print(f"# {os.path.relpath(lib, ROOT)}: {len(names)} kernels (sm_100a); per-kernel counts of Blackwell mnemonics")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
tot = collections.Counter()
for n in sorted(names, key=short):
    c = counts[n]
    tot.update(c)
    if not any(c[m] for m in MNEMONICS[:10]):
        continue
    print(f"{short(n):58s} {c['_total']:7d} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
print(f"{'TOTAL (all kernels incl. those without these mnemonics)':58s} {tot['_total']:7d} " +
      " ".join(f"{tot[m]:8d}" for m in MNEMONICS))
