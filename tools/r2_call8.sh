#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c8_pytest.log 2>&1
tail -3 $O/r2c8_pytest.log
DMB_TRACE=1 python tools/gpu_slow_step_probe.py 4096 110 > $O/r2c8_slow_probe.txt 2>&1
python - <<'PY'
import re
sp=[float(m.group(1)) for m in re.finditer(r"span\s+([0-9.]+) us", open("gpurun_out/r2c8_slow_probe.txt").read())]
import statistics as st
print("spans: n", len(sp), "mean", round(st.mean(sp),1), "median", round(st.median(sp),1), "min", min(sp), "max", max(sp))
PY
grep "> 32: [1-9]" $O/r2c8_slow_probe.txt | head -5 | cut -c1-250
python bench.py --no-cpu-baseline > $O/r2c8_bench_1gpu.json 2>/dev/null; cut -c1-250 $O/r2c8_bench_1gpu.json
for c in 3 4 5; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > $O/r2c8_bench_config$c.json 2>/dev/null; cut -c60-200 $O/r2c8_bench_config$c.json; done
