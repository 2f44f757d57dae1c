#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29566"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), d['config']['launch']['grid'], d.get('rank_spread'))" "$1"; }
$TR bench.py --gpus 8 --steps 200 --warmup 20 2>/dev/null | tee $O/r2c22_config2_8gpu.json | pr "8gpu ctas1 reserve1"
DMB_RESERVE_SMS=0 $TR bench.py --gpus 8 --steps 200 --warmup 20 --nccl-ctas 0 2>/dev/null | pr "8gpu default reserve0"
$TR bench.py --gpus 8 --steps 200 --warmup 20 --nccl-ctas 4 2>/dev/null | pr "8gpu ctas4 reserve1"
DMB_RESERVE_SMS=0 $TR bench.py --gpus 8 --steps 100 --warmup 10 --config 4 --nccl-ctas 0 2>/dev/null | pr "8gpu config4 default reserve0"
$TR bench.py --gpus 8 --steps 100 --warmup 10 --config 4 2>/dev/null | tee $O/r2c22_config4_8gpu.json | pr "8gpu config4 ctas1 reserve1"
