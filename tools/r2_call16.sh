#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c16_pytest.log 2>&1; tail -2 $O/r2c16_pytest.log
bash tools/ab_bench.sh 2 default pinS 2>&1 | tee $O/r2c16_ab.txt
for kn in "DMB_SPREAD=2" "DMB_SPREAD=1" "DMB_COST_MODE=1" "DMB_COST_MODE=2"; do
  echo "== $kn"; env $kn python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
