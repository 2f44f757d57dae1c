#!/bin/bash
# Round artefacts on one B200: GPU test-suite, bench (both arms), ncu launch list + full capture of k_step,
# phase timers, CTA trace.  Everything lands in gpurun_out/final_*.
set -u
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -m gpu -q) > $O/final_pytest.log 2>&1
python bench.py > $O/final_bench_1gpu.json 2> $O/final_bench_1gpu.err
# (reference arm: CPU only, see profiles/r1_v12_bench_reference_arm.json)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 60 --warmup 20 --no-cpu-baseline > $O/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/final_kstep \
    python bench.py --steps 80 --warmup 20 --no-cpu-baseline > $O/final_ncu.log 2>&1
DMB_TRACE=1 python tools/gpu_cta_trace.py 4096 > $O/final_cta_trace.txt 2>&1
if [ -f variants/libdmb200_timers.so ]; then
  DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/final_phase_timers.txt 2>&1
fi
python tools/gpu_rollout_bench.py > $O/final_rollout_policy.txt 2>&1
tail -3 $O/final_pytest.log; cut -c1-400 $O/final_bench_1gpu.json; 
