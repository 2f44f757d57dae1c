#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c17_pytest.log 2>&1; tail -3 $O/r2c17_pytest.log
python tools/gpu_rollout_bench.py 2>&1 | tail -3 | tee $O/r2c17_rollout_policy.txt
DMB_LIB=$PWD/variants/libdmb200_blk2d.so python -m pytest tests -m gpu -q -k "forward or step or many or reset or odd" > $O/r2c17_pytest_blk2d.log 2>&1; tail -2 $O/r2c17_pytest_blk2d.log
bash tools/ab_bench.sh 2 default blk2d 2>&1 | tee $O/r2c17_ab.txt
