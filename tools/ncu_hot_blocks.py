#!/usr/bin/env python3
"""Hot straight-line SASS blocks of an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass):
contiguous instructions with the same execution count, ranked by their share of all executed instructions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
envs = float(sys.argv[3]) if len(sys.argv) > 3 else 4096.0
cur = hdr = line = None
blocks = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iA = hdr.index("Address"); iS = hdr.index("Source"); continue
    if hdr is None:
        continue
    if r[0].isdigit():
        line = (cur, int(r[0])); continue
    if len(r) > iA and r[iA].startswith('0x'):
        try:
            n = int(r[iI])
        except ValueError:
            n = 0
        blocks.append((int(r[iA], 16), line, r[iS].strip(), n))
blocks.sort()
tot = sum(b[3] for b in blocks)
groups, g = [], []
for b in blocks:
    if g and b[3] == g[-1][3] and b[0] == g[-1][0] + 16:
        g.append(b)
    else:
        if g:
            groups.append(g)
        g = [b]
groups.append(g)
groups.sort(key=lambda g: -len(g) * g[0][3])
print("total executed warp-instructions", tot, "static", len(blocks))
for g in groups[:topn]:
    n, L = g[0][3], len(g)
    srcs = collections.Counter(b[1] for b in g)
    top = ", ".join(f"{k[0].replace('dmb_', '').replace('.cuh', '')}:{k[1]}x{v}" for k, v in srcs.most_common(4))
    print(f"{100 * n * L / tot:5.2f}%  exec/env {n / envs:7.1f}  rows {L:4d}  {top}")
