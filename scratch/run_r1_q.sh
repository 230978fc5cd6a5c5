#!/bin/bash
# tests + final-size bench + ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_r1_q.err > gpurun_out/bench_r1_q.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1_q.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_r1_q.err
