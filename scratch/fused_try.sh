#!/bin/bash
cd /root/repo

for cfg in "2368 1" "1184 1"; do
  set -- $cfg
  echo "== XB_TILE=$1 XB_FUSED=$2"
  XB_TILE=$1 XB_FUSED=$2 timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/fused.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'ft_call',round(d['formTangent_call_ms'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.01}, 'launches', d['gpu_launches'])" || tail -5 gpurun_out/fused.err
done
