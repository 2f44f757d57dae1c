#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c6_pytest.log 2>&1
tail -3 $O/r2c6_pytest.log
DMB_TRACE=1 python tools/gpu_time_probe.py 4096 2>&1 | tee $O/r2c6_time_probe_trace.txt
python bench.py --no-cpu-baseline > $O/r2c6_bench_1gpu.json 2>/dev/null; cut -c1-250 $O/r2c6_bench_1gpu.json
for c in 3 5; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > $O/r2c6_bench_config$c.json 2>/dev/null; cut -c1-250 $O/r2c6_bench_config$c.json; done
python bench.py --config 4 --steps 100 --warmup 10 --no-cpu-baseline > $O/r2c6_bench_config4_1gpu.json 2>/dev/null; cut -c1-250 $O/r2c6_bench_config4_1gpu.json
