#!/bin/bash
# kernel_ms of the brick step for each tangent kernel variant (n^3 elements)
cd /root/repo
N=${1:-128}
for v in 0 1 2 3; do
  python bench.py --n $N --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --opt tangent_variant=$v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant $v', 'ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items() if x>0.01})
"
done
