#!/usr/bin/env python3
"""Aggregate an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass) by CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
inst = collections.Counter(); samp = collections.Counter(); text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or r[0] == "" or not r[0].isdigit():
        continue
    try:
        n, w = int(r[iI]), int(r[iW])
    except ValueError:
        continue
    k = (cur_file, int(r[0]))
    inst[k] += n; samp[k] += w; text[k] = r[1].strip()[:100]
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-instructions", ti, "stall samples", ts)
print("---- by stall samples")
for k, v in samp.most_common(topn):
    print(f"{k[0]}:{k[1]:4d} samp {100*v/ts:5.1f}%  inst {100*inst[k]/ti:5.1f}% | {text[k]}")
