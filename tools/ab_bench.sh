#!/bin/bash
# A/B the in-tree kernel builds on the GPU box: default library + every variants/libdmb200_*.so given by name.
#   tools/ab_bench.sh [reps] name1 name2 ...      (name "default" = deepmimic_mujoco_b200/libdmb200.so)
reps=${1:-2}; shift
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
  for n in "$@"; do
    if [ "$n" = "default" ]; then lib=""; else lib="$PWD/variants/libdmb200_$n.so"; fi
    v=$(DMB_LIB=$lib python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/ab_$n.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4fM  e2e %.4fM  %.4f ms' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))")
    echo "$n rep$r: $v"
  done
done
