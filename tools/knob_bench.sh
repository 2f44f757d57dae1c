#!/bin/bash
# bench the default library under environment-variable knobs:  tools/knob_bench.sh "DMB_ENVS_PER_CTA=13" "DMB_COST_MODE=2" ...
for k in "$@"; do
  v=$(env $k python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4fM  e2e %.4fM  %.4f ms' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))")
  echo "$k: $v"
done
