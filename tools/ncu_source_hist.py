#!/usr/bin/env python
"""Per-opcode and per-region executed-instruction histogram from `ncu --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_source_hist.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
iS, iE, iSm = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
data = []
for r in rows[hdr + 1:]:
    if len(r) <= iE:
        continue
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.rstrip(";")
    e = int(r[iE] or 0); s = int(r[iSm] or 0)
    ops[op.split(".")[0]] += e; samp[op.split(".")[0]] += s; tot += e; tots += s
    data.append((e, s, src))
print(f"total warp instructions {tot:,}  samples {tots:,}")
for op, e in ops.most_common(24):
    print(f"  {op:14s} {e:14,d} {100 * e / tot:6.2f}%   stall-samples {100 * samp[op] / max(tots, 1):6.2f}%")
print("top stall-sample instructions:")
for e, s, src in sorted(data, key=lambda x: -x[1])[:16]:
    print(f"  {s:7d} samples  {e:12,d} exec   {src[:90]}")
