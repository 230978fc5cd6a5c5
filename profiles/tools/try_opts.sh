#!/bin/bash
# kernel_ms of the brick step for option sets: profiles/tools/try_opts.sh N lib "opt1=v,opt2=v" ...
cd /root/repo
N=$1; LIB=$2; shift; shift
for o in "$@"; do
  args=""; for kv in ${o//,/ }; do args="$args --opt $kv"; done
  XARA_B200_LIB=$PWD/xara_b200/$LIB python bench.py --n $N --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$LIB $o', 'ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items() if x>0.01})
"
done
