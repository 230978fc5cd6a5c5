#!/bin/bash
# bench.py under torchrun at N = $1 (the driver's own launch line)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_r1d_n$N.err > gpurun_out/bench_r1d_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1d_n$N.json'))
print('N=$N ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],2))
PY
tail -2 gpurun_out/bench_r1d_n$N.err | cut -c1-300
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "nccl" 2>&1 | tail -3; fi
