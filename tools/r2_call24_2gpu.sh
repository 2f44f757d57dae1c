#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588"
timeout 200 $TR tools/gpu_peer_gather_check.py 1000 2>&1 | tail -4 | tee $O/r2c24_peer_check.txt
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2), d.get('rank_spread'), d['config'].get('p2p_gather_unavailable'))" "$1"; }
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 20 2>$O/r2c24_p2p.err | tee $O/r2c24_bench_2gpu_p2p.json | pr "2gpu p2p"
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 20 --gather nccl 2>/dev/null | pr "2gpu nccl"
tail -3 $O/r2c24_p2p.err
