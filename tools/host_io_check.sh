#!/bin/bash
# One short call: GPU tests (incl. tests/test_gpu_host_io.py), the default bench line (every e2e form: memcpy /
# host-mapped action / host-mapped record), configs 3 and 5 on one GPU, smoke.  Results: gpurun_out/hostio_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/hostio_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/hostio_pytest.log
tail -5 gpurun_out/hostio_pytest.log
python bench.py --steps 100 --warmup 5 > gpurun_out/hostio_bench.json 2> gpurun_out/hostio_bench.err; echo "bench rc=$?"
cat gpurun_out/hostio_bench.json
for c in 3 5; do
  python bench.py --config $c --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/hostio_bench_config$c.json 2> gpurun_out/hostio_bench_config$c.err; echo "bench config $c rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/hostio_bench_config$c.json')); print($c, d['value'], d['e2e'])"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
