#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2), d.get('rank_spread'), d['config'].get('p2p_gather_unavailable'))" "$1"; }
timeout 90 $TR bench.py --gpus 8 --steps 200 --warmup 20 2>$O/fused_p2p.err | tee $O/fused_config2_8gpu_p2p.json | pr "8gpu config2 p2p"
timeout 90 $TR bench.py --gpus 8 --steps 200 --warmup 20 --gather nccl 2>/dev/null | tee $O/fused_config2_8gpu_nccl.json | pr "8gpu config2 nccl"
timeout 90 $TR bench.py --gpus 8 --steps 100 --warmup 10 --config 4 2>/dev/null | tee $O/fused_config4_8gpu_p2p.json | pr "8gpu config4 p2p"
timeout 90 $TR bench.py --gpus 8 --steps 100 --warmup 10 --config 5 2>/dev/null | tee $O/fused_config5_8gpu_p2p.json | pr "8gpu config5 p2p"
timeout 90 $TR tools/gpu_peer_gather_check.py 512 2>&1 | tail -2 | tee $O/fused_peer_check_8.txt
tail -3 $O/fused_p2p.err
