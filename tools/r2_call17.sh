#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c17_pytest.log 2>&1; tail -3 $O/r2c17_pytest.log
python tools/gpu_rollout_bench.py 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 100 2>/dev/null | cut -c1-200
