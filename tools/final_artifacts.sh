#!/bin/bash
# Round-2 evidence from one B200: GPU tests, bench (both arms), sanitizer runs, ncu launch list + full capture of the
# shipped k_step, phase timers, CTA trace, slow-step probe, policy rollout.  Everything lands in gpurun_out/r2f_*.
set -u
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -m gpu -q) > $O/r2f_pytest.log 2>&1; tail -3 $O/r2f_pytest.log
cp $O/parity_measured.json $O/r2f_parity_measured.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.txt 2>&1; tail -1 $O/r2f_smoke.txt
python bench.py > $O/r2f_bench_1gpu.json 2> $O/r2f_bench_1gpu.err; cut -c1-260 $O/r2f_bench_1gpu.json
python bench.py --impl reference --steps 10 --warmup 2 > $O/r2f_bench_reference.json 2>/dev/null; cut -c1-160 $O/r2f_bench_reference.json
for c in 3 4 5; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > $O/r2f_bench_config${c}_1gpu.json 2>/dev/null; cut -c60-200 $O/r2f_bench_config${c}_1gpu.json; done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_small_rollout.py 64 6 > $O/r2f_memcheck.log 2>&1; tail -2 $O/r2f_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_small_rollout.py 64 3 > $O/r2f_racecheck.log 2>&1; tail -3 $O/r2f_racecheck.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file $O/r2f_launches.csv \
    python bench.py --steps 60 --warmup 20 --no-cpu-baseline > $O/r2f_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/r2f_kstep \
    python bench.py --steps 80 --warmup 20 --no-cpu-baseline > $O/r2f_ncu.log 2>&1
DMB_TRACE=1 python tools/gpu_cta_trace.py 4096 > $O/r2f_cta_trace.txt 2>&1; tail -4 $O/r2f_cta_trace.txt | cut -c1-200
DMB_TRACE=1 python tools/gpu_slow_step_probe.py 4096 110 > $O/r2f_slow_probe.txt 2>&1
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2f_phase_timers.txt 2>&1
python tools/gpu_rollout_bench.py > $O/r2f_rollout_policy.txt 2>&1; tail -2 $O/r2f_rollout_policy.txt
python tools/gpu_latency_scan.py > $O/r2f_latency_scan.txt 2>&1; cat $O/r2f_latency_scan.txt
