#!/bin/bash
set -u
O=gpurun_out
for kn in "DMB_SYNC_MASK=0x41" "DMB_SYNC_MASK=0x00" "DMB_SYNC_MASK=0x02" "DMB_SYNC_MASK=0x08" "DMB_SYNC_MASK=0x10" "DMB_SYNC_MASK=0x09"; do
  echo "== $kn"; env $kn python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c12_phase_timers.txt 2>&1
cat $O/r2c12_phase_timers.txt
DMB_SYNC_MASK=0x00 DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c12_phase_timers_mask0.txt 2>&1
cat $O/r2c12_phase_timers_mask0.txt
