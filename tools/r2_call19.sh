#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c19_pytest.log 2>&1; tail -2 $O/r2c19_pytest.log
DMB_PARITY_REPORT=$O/r2c19_parity_fastdiv.json DMB_LIB=$PWD/variants/libdmb200_fastdiv.so python -m pytest tests -m gpu -q > $O/r2c19_pytest_fastdiv.log 2>&1; tail -2 $O/r2c19_pytest_fastdiv.log
bash tools/ab_bench.sh 3 default fastdiv 2>&1 | tee $O/r2c19_ab.txt
