#!/bin/bash
# Round-2 call 4: A/B of register-pinning / inlining variants, corrected CTA trace (true CTA end + diagnostics), knobs
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/r2c4_pytest.log 2>&1
tail -2 $O/r2c4_pytest.log
bash tools/ab_bench.sh 2 default nopin inlkin 2>&1 | tee $O/r2c4_ab.txt
DMB_TRACE=1 TRACE_FLUSH=1 python tools/gpu_cta_trace.py 4096 > $O/r2c4_cta_trace_cold.txt 2>&1
tail -12 $O/r2c4_cta_trace_cold.txt
for kn in "DMB_PATIENCE=3000" "DMB_PATIENCE=10000" "DMB_SYNC_MASK=0x43" "DMB_SYNC_MASK=0x49" "DMB_SYNC_MASK=0x51" "DMB_COST_MODE=2" "DMB_NO_SORT=1"; do
  echo "== $kn" | tee -a $O/r2c4_knobs.txt
  env $kn python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" 2>&1 | tee -a $O/r2c4_knobs.txt
done
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c4_phase_timers.txt 2>&1
cat $O/r2c4_phase_timers.txt
