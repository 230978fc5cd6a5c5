#!/bin/bash
cd /root/repo
for lib in libxara_b200_v0.so libxara_b200.so; do
echo "== $lib"
XARA_B200_LIB=/root/repo/xara_b200/$lib python bench.py --n 128 --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/bq.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'ft_call',round(d['formTangent_call_ms'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.01}, 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bq.err
done
