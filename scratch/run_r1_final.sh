#!/bin/bash
# round-end evidence: tests, headline bench, reference arm, ncu launch list of the same command, ncu --set full of the step's kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_r1_final2.err > gpurun_out/bench_r1_final2.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1_final2.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_r1_final2_ref.err > gpurun_out/bench_r1_final2_ref.json; cut -c1-300 gpurun_out/bench_r1_final2_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch_final2.log 2>&1
tail -1 gpurun_out/ncu_launch_final2.log | cut -c1-200
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'brick_tangent_sym|brick_update|assemble_A|assemble_B' -s 8 -c 4 -o gpurun_out/prof_r1_final2 python bench.py --n 96 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_final2.log 2>&1
tail -2 gpurun_out/ncu_full_final2.log | cut -c1-200
