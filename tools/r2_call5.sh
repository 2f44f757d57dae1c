#!/bin/bash
set -u
O=gpurun_out
DMB_TRACE=1 python tools/gpu_time_probe.py 4096 2>&1 | tee $O/r2c5_time_probe_trace.txt
DMB_TRACE=0 python tools/gpu_time_probe.py 4096 2>&1 | tee $O/r2c5_time_probe_notrace.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file $O/r2c5_launches.csv \
    python bench.py --steps 60 --warmup 20 --no-cpu-baseline > $O/r2c5_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2c5_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:60s} n={len(v):4d} mean {sum(v)/len(v):10.1f} max {max(v):10.1f}")
PY
