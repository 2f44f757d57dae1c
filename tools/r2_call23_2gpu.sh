#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
timeout 300 $TR tools/gpu_peer_gather_check.py 1000 2>&1 | tail -8 | tee $O/r2c23_peer_check.txt
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2), d.get('rank_spread'), d['config'].get('p2p_gather_unavailable'))" "$1"; }
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | pr 1gpu
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 2>$O/r2c23_p2p.err | tee $O/r2c23_bench_2gpu_p2p.json | pr "2gpu p2p"
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 --gather nccl 2>/dev/null | pr "2gpu nccl"
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 --config 4 2>/dev/null | pr "2gpu config4 p2p"
tail -5 $O/r2c23_p2p.err
python -m pytest tests -m gpu -q 2>&1 | tail -2
