#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), d['config']['launch']['grid'])" "$1"; }
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | pr 1gpu
$TR bench.py --gpus 2 --steps 200 --warmup 20 2>/dev/null | pr "2gpu ctas1 reserve1"
$TR bench.py --gpus 2 --steps 200 --warmup 20 --nccl-ctas 0 2>/dev/null | pr "2gpu ctas-default reserve1"
DMB_RESERVE_SMS=0 $TR bench.py --gpus 2 --steps 200 --warmup 20 --nccl-ctas 0 2>/dev/null | pr "2gpu ctas-default reserve0"
$TR bench.py --gpus 2 --steps 100 --warmup 10 --config 4 2>/dev/null | pr "2gpu config4 ctas1 reserve1"
DMB_RESERVE_SMS=0 $TR bench.py --gpus 2 --steps 100 --warmup 10 --config 4 --nccl-ctas 0 2>/dev/null | pr "2gpu config4 default reserve0"
