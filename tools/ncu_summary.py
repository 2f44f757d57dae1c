#!/usr/bin/env python3
"""Summarise an ncu source-page CSV: stall reasons, per-function samples/instructions (dynamic and static)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
dev = open(sys.argv[2]).read().split('\n') if len(sys.argv) > 2 else []
starts = []
for i, l in enumerate(dev, 1):
    m = re.match(r'__device__ .*? (\w+)\(', l)
    if m: starts.append((i, m.group(1)))
def fn_of(line):
    name = '?'
    for s, n in starts:
        if s <= line: name = n
    return name
cur = hdr = None; key = None
inst = collections.Counter(); samp = collections.Counter(); static = collections.Counter(); stalls = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); iA = hdr.index("Address")
        sidx = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}; continue
    if hdr is None: continue
    if r[0].isdigit():
        key = fn_of(int(r[0])) if cur == 'dmb_device.cuh' else cur
        try: inst[key] += int(r[iI]); samp[key] += int(r[iW])
        except ValueError: pass
        for h, i in sidx.items():
            try: stalls[h] += int(r[i])
            except ValueError: pass
    elif len(r) > iA and r[iA].startswith('0x') and key: static[key] += 1
ti, ts = sum(inst.values()), sum(samp.values()); ss = sum(stalls.values())
print("total warp-instr", ti, "samples", ts)
print("stalls:", ", ".join(f"{k[6:]} {100*v/ss:.1f}%" for k, v in stalls.most_common(8)))
for k, v in samp.most_common(22):
    print(f"{k:22s} samples {100*v/ts:5.1f}%  inst {100*inst[k]/ti:5.1f}%  static-sass-rows(shown) {static[k]}")
