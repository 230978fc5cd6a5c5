#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_r1e.err > gpurun_out/bench_r1e.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1e.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['clocks'])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
