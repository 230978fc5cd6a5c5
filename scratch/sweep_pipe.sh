#!/bin/bash
cd /root/repo
run() {
  echo "== $*"
  env "$@" python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/sweep_pipe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'ft_call',round(d['formTangent_call_ms'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.01}, 'launches', d['gpu_launches'])" || tail -5 gpurun_out/sweep_pipe.err
}
run XB_TILE=0 XB_ASM_CARVEOUT=1
run XB_TILE=0 XB_PIPELINE=8 XB_TAN_WARPS=4 XB_ASM_CARVEOUT=1
run XB_TILE=0 XB_PIPELINE=8 XB_TAN_WARPS=3 XB_ASM_CARVEOUT=1
run XB_TILE=0 XB_PIPELINE=32 XB_TAN_WARPS=3 XB_ASM_CARVEOUT=1
run XB_TILE=37888 XB_AHEAD=4 XB_TAN_WARPS=3 XB_ASM_CARVEOUT=1
