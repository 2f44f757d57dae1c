#!/bin/bash
# 8-GPU scaling table with the final binary (configs 2 / 4 / 5; gather depth 4 and 2)
set -u
O=gpurun_out
run() { # nproc, tag, args...
  n=$1; tag=$2; shift 2
  if [ $n -eq 1 ]; then python bench.py --gpus 1 "$@" > $O/scale_$tag.json 2> $O/scale_$tag.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@" > $O/scale_$tag.json 2> $O/scale_$tag.err; fi
  python - "$O/scale_$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(f"{sys.argv[2]:34s} {d['value']/1e6:8.2f} M  {d['ms_per_step']:.4f} ms  e2e {d['e2e']['value']/1e6:7.2f} M  n_gpus {d['n_gpus']} envs {d['config']['envs_global']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run 1 config2_1gpu --steps 200 --warmup 20 --no-cpu-baseline
run 2 config2_weak_2gpu --steps 200 --warmup 20
run 4 config2_weak_4gpu --steps 200 --warmup 20
run 8 config2_weak_8gpu --steps 200 --warmup 20
run 8 config2_weak_8gpu_depth2 --steps 200 --warmup 20 --gather-depth 2
run 8 config4_weak_8gpu --config 4 --steps 100 --warmup 10
run 4 config4_strong_4gpu --config 4 --scaling strong --envs-global 65536 --steps 100 --warmup 10
run 2 config4_strong_2gpu --config 4 --scaling strong --envs-global 65536 --steps 60 --warmup 10
run 1 config4_strong_1gpu --config 4 --scaling strong --envs-global 65536 --steps 40 --warmup 5 --no-cpu-baseline
run 8 config5_weak_8gpu --config 5 --steps 100 --warmup 10
