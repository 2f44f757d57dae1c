#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c10_pytest.log 2>&1
tail -2 $O/r2c10_pytest.log
python bench.py > $O/r2c10_bench_1gpu.json 2>$O/r2c10_bench_1gpu.err; cut -c1-300 $O/r2c10_bench_1gpu.json
for kn in "DMB_SORT_PERIOD=1" "DMB_SORT_PERIOD=4" "DMB_SORT_PERIOD=32"; do
  echo "== $kn"; env $kn python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
done
DMB_TRACE=1 python tools/gpu_slow_step_probe.py 4096 110 > $O/r2c10_slow_probe.txt 2>&1
python - <<'PY'
import re, statistics as st
sp=[float(m.group(1)) for m in re.finditer(r"span\s+([0-9.]+) us", open("gpurun_out/r2c10_slow_probe.txt").read())]
print("spans: n", len(sp), "mean", round(st.mean(sp),1), "median", round(st.median(sp),1), "min", min(sp), "max", max(sp))
PY
ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/r2c10_kstep \
    python bench.py --steps 80 --warmup 20 --no-cpu-baseline > $O/r2c10_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file $O/r2c10_launches.csv \
    python bench.py --steps 60 --warmup 20 --no-cpu-baseline > $O/r2c10_launches.log 2>&1
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c10_phase_timers.txt 2>&1
cat $O/r2c10_phase_timers.txt
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2c10_bench_reference.json 2>/dev/null; cut -c1-200 $O/r2c10_bench_reference.json
