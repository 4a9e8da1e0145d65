#!/usr/bin/env python
"""Warp-stall reasons summed over a kernel from `ncu --page source --csv` output, overall and per opcode.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
iS = H.index("Source")
tot = collections.Counter(); per_op = collections.defaultdict(collections.Counter)
for r in rows[hdr + 1:]:
    if len(r) <= max(cols):
        continue
    toks = r[iS].split()
    op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
    for i in cols:
        v = int(r[i] or 0)
        tot[H[i]] += v; per_op[op][H[i]] += v
n = sum(tot.values())
print("all samples", n)
for k, v in tot.most_common():
    print(f"  {k:24s} {v:8d} {100 * v / n:6.2f}%")
for op, c in sorted(per_op.items(), key=lambda x: -sum(x[1].values()))[:12]:
    s = sum(c.values())
    print(f"{op:8s} {100 * s / n:5.1f}%  " + "  ".join(f"{k[6:]}={100 * v / s:.0f}%" for k, v in c.most_common(4)))
