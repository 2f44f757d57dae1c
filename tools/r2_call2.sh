#!/bin/bash
# Round-2 call 2: first run of the 28-envs-per-SM tile kernel: GPU tests, bench, knobs, latency scan
set -u
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -m gpu -q -s) > $O/r2c2_pytest.log 2>&1
cp $O/parity_measured.json $O/r2c2_parity_measured.json 2>/dev/null
python bench.py --no-cpu-baseline > $O/r2c2_bench_1gpu.json 2> $O/r2c2_bench_1gpu.err
for kn in "DMB_SPREAD=0" "DMB_GROUPS=2" "DMB_GROUPS=4" "DMB_ENVS_PER_CTA=14" "DMB_ENVS_PER_CTA=21" "DMB_SYNC_MASK=0x01" "DMB_SYNC_MASK=0x7f" "DMB_LOCKSTEP=0"; do
  echo "== $kn" >> $O/r2c2_knobs.txt
  env $kn python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['launch'])" >> $O/r2c2_knobs.txt 2>&1
done
for E in 8192 16384; do
  echo "== envs $E" >> $O/r2c2_knobs.txt
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --envs-per-gpu $E 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['launch'])" >> $O/r2c2_knobs.txt 2>&1
done
DMB_TRACE=1 python tools/gpu_cta_trace.py 4096 > $O/r2c2_cta_trace.txt 2>&1
tail -15 $O/r2c2_pytest.log; cut -c1-400 $O/r2c2_bench_1gpu.json; cat $O/r2c2_knobs.txt; tail -8 $O/r2c2_cta_trace.txt
