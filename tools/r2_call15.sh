#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c15_pytest_pair17.log 2>&1; tail -2 $O/r2c15_pytest_pair17.log
DMB_LIB=$PWD/variants/libdmb200_pair9.so python -m pytest tests -m gpu -q -k "forward or step or many or full_size" > $O/r2c15_pytest_pair9.log 2>&1; tail -2 $O/r2c15_pytest_pair9.log
bash tools/ab_bench.sh 2 default pair9 nopair 2>&1 | tee $O/r2c15_ab.txt
DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py 4096 > $O/r2c15_phase_timers.txt 2>&1
cat $O/r2c15_phase_timers.txt
