#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in quad frame frame3d; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bm.err > gpurun_out/bench_r1e_$wl.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r1e_$wl.json'))
print('$wl ms/step',round(d['ms_per_step'],3),'kernels',{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.005}, 'e2e', round(d['e2e']['ms_per_step'],3))" || tail -5 gpurun_out/bm.err
done
