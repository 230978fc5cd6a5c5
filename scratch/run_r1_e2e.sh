#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --n 160 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r1_p.err > gpurun_out/bench_r1_p.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1_p.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'])
PY
