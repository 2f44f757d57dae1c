#!/bin/bash
# One short call: GPU tests (incl. tests/test_gpu_host_io.py), smoke, the default bench line (e2e through the
# host-mapped step next to the memcpy form).  Results: gpurun_out/hostio_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/hostio_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/hostio_pytest.log
tail -5 gpurun_out/hostio_pytest.log
python bench.py --steps 100 --warmup 5 > gpurun_out/hostio_bench.json 2> gpurun_out/hostio_bench.err; echo "bench rc=$?"
cat gpurun_out/hostio_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
