#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q -k "tiled" 2>&1 | tail -5
for cfg in "0 0" "9472 1" "4736 1" "18944 1"; do
  set -- $cfg
  echo "== XB_TILE=$1 XB_FUSED=$2"
  XB_TILE=$1 XB_FUSED=$2 timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/fused.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'ft_call',round(d['formTangent_call_ms'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.01}, 'launches', d['gpu_launches'])" || tail -5 gpurun_out/fused.err
done
