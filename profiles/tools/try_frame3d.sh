#!/bin/bash
cd /root/repo
for lib in "$@"; do
  XARA_B200_LIB=$PWD/xara_b200/$lib python bench.py --workload ${WL:-frame3d} --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items() if x>0.01})
"
done
