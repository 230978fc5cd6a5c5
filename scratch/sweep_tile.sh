#!/bin/bash
# sweep of the tiled formTangent pipeline on the 160^3 block
cd /root/repo

for cfg in "9472 0" "4736 0" "18944 0" "9472 6" "4736 8" "37888 0"; do
  set -- $cfg
  echo "== XB_TILE=$1 XB_AHEAD=$2"
  XB_TILE=$1 XB_AHEAD=$2 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/sweep_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'ft_call',round(d['formTangent_call_ms'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.01},'setup',d['config']['setup_s'], 'launches', d['gpu_launches'])"
done
