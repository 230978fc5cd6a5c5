#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload quad --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>gpurun_out/bq.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('quad ms/step',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['kernel_ms'].items() if v>0.005})"
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_r1e.err > gpurun_out/bench_r1e.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1e.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['clocks'])
PY
