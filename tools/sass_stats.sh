#!/bin/bash
# (build container, no GPU) static record of the shipped library: ptxas resource usage per kernel and the SASS
# instruction mix of the step kernel.  usage: tools/sass_stats.sh > profiles/<round>_sass_stats.txt
set -e
cd "$(dirname "$0")/.."
SO=deepmimic_mujoco_b200/libdmb200.so
echo "# $(sha256sum $SO | cut -c1-16)  $SO  ($(stat -c %s $SO) bytes)"
echo "# --- cuobjdump -res-usage"
cuobjdump -res-usage $SO 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/ Fatbin.*//' | awk '{$1=$1; print}' | cut -c1-260
echo "# --- SASS of k_step<true> (the lockstep step kernel): instructions by mnemonic"
cuobjdump -sass -fun '_Z6k_stepILb1EEv7DevPtrs9dmb_statePKf12dmb_step_outiyj' $SO 2>/dev/null > /tmp/kstep.sass || true
if [ ! -s /tmp/kstep.sass ] || ! grep -q "/\*0" /tmp/kstep.sass; then cuobjdump -sass $SO > /tmp/all.sass; awk '/Function : .*k_stepILb1/{f=1} /Function : /{if(!/k_stepILb1/)f=0} f' /tmp/all.sass > /tmp/kstep.sass; fi
grep -E "^\s+/\*[0-9a-f]{4,}\*/" /tmp/kstep.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{split($1,a,"."); print a[1]}' > /tmp/kstep.mn
echo "total $(wc -l < /tmp/kstep.mn) instructions"
sort /tmp/kstep.mn | uniq -c | sort -k1,1nr | head -48
