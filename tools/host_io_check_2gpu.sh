#!/bin/bash
# (2-GPU box) the default bench line on 2 ranks: every e2e form together with the fused peer all-gather.
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/hostio_bench_2gpu.json 2> gpurun_out/hostio_bench_2gpu.err; echo "bench rc=$?"
cat gpurun_out/hostio_bench_2gpu.json; tail -3 gpurun_out/hostio_bench_2gpu.err
