#!/bin/bash
# correctness with the new default tangent kernel, then A/B bench, then one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/sym_tests.log
cat gpurun_out/sym_tests.log
for v in sym tile; do
  XB_TANGENT=$v timeout 600 python bench.py --n 160 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/sym_bench_$v.err > gpurun_out/sym_bench_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/sym_bench_$v.json'))
print('$v', 'ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()})
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'brick_tangent_sym|brick_update' -c 2 -o gpurun_out/prof_r1_sym python bench.py --n 96 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_sym.log 2>&1
tail -3 gpurun_out/ncu_sym.log
