#!/bin/bash
set -u
O=gpurun_out
DMB_LIB=$PWD/variants/libdmb200_blk2d.so python -m pytest tests -m gpu -q > $O/r2c18_pytest_blk2d.log 2>&1; tail -2 $O/r2c18_pytest_blk2d.log
bash tools/ab_bench.sh 3 default blk2d 2>&1 | tee $O/r2c18_ab.txt
