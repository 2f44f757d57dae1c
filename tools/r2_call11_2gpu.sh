#!/bin/bash
# 2-GPU run: NCCL all-gather overlapped with the next step vs on the compute stream; TRPO 2-rank sync on NCCL
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 200 --warmup 20 > $O/r2c11_bench_2gpu.json 2> $O/r2c11_bench_2gpu.err; cut -c1-300 $O/r2c11_bench_2gpu.json
$TR bench.py --gpus 2 --steps 200 --warmup 20 --sync-gather > $O/r2c11_bench_2gpu_syncgather.json 2>/dev/null; cut -c60-200 $O/r2c11_bench_2gpu_syncgather.json
$TR bench.py --gpus 2 --steps 100 --warmup 10 --config 4 > $O/r2c11_bench_config4_2gpu.json 2>/dev/null; cut -c60-200 $O/r2c11_bench_config4_2gpu.json
$TR bench.py --gpus 2 --steps 100 --warmup 10 --config 4 --scaling strong --envs-global 65536 > $O/r2c11_bench_config4_strong_2gpu.json 2>/dev/null; cut -c60-200 $O/r2c11_bench_config4_strong_2gpu.json
python bench.py --gpus 1 --steps 60 --warmup 10 --config 4 --scaling strong --envs-global 65536 --no-cpu-baseline > $O/r2c11_bench_config4_strong_1gpu.json 2>/dev/null; cut -c60-200 $O/r2c11_bench_config4_strong_1gpu.json
$TR tools/train_trpo.py --envs 1024 --horizon 16 --iters 3 > $O/r2c11_trpo_2rank.txt 2>&1; tail -4 $O/r2c11_trpo_2rank.txt
