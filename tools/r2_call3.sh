#!/bin/bash
# Round-2 call 3: diagnostics of the 28-env kernel: tests, cold-start probe, traces, phase timers, ncu full
set -u
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -m gpu -q -s) > $O/r2c3_pytest.log 2>&1
cp $O/parity_measured.json $O/r2c3_parity_measured.json 2>/dev/null
python tools/gpu_cold_start_probe.py 4096 > $O/r2c3_cold_start.txt 2>&1
python bench.py --no-cpu-baseline --no-flush > $O/r2c3_bench_noflush.json 2> /dev/null
DMB_TRACE=1 python tools/gpu_cta_trace.py 4096 > $O/r2c3_cta_trace_warm.txt 2>&1
DMB_TRACE=1 TRACE_FLUSH=1 python tools/gpu_cta_trace.py 4096 > $O/r2c3_cta_trace_cold.txt 2>&1
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c3_phase_timers.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/r2c3_kstep \
    python bench.py --steps 80 --warmup 20 --no-cpu-baseline > $O/r2c3_ncu.log 2>&1
tail -4 $O/r2c3_pytest.log; cat $O/r2c3_cold_start.txt; cut -c1-200 $O/r2c3_bench_noflush.json; tail -3 $O/r2c3_cta_trace_warm.txt; tail -3 $O/r2c3_cta_trace_cold.txt; cat $O/r2c3_phase_timers.txt
