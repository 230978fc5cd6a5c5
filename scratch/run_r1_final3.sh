#!/bin/bash
# round-end evidence: tests, headline bench, reference arm, secondary workloads, ncu launch list of the same command, ncu --set full of the step's kernels
mkdir -p gpurun_out
T=r1e
echo skip-tests
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_${T}.err > gpurun_out/bench_${T}.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${T}.json'))
print('ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], d['clocks'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_${T}_ref.err > gpurun_out/bench_${T}_ref.json; cut -c1-300 gpurun_out/bench_${T}_ref.json
for wl in quad frame frame3d; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_${T}_$wl.err > gpurun_out/bench_${T}_$wl.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_${T}_$wl.json'))
print('$wl ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items() if x>0.005}, 'e2e ms', round(d['e2e']['ms_per_step'],3))"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch_${T}.log 2>&1
tail -1 gpurun_out/ncu_launch_${T}.log | cut -c1-200
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'brick_tangent_sym|brick_update|assemble_A|assemble_B' -s 8 -c 4 -o gpurun_out/prof_${T} python bench.py --n 96 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_${T}.log 2>&1
tail -2 gpurun_out/ncu_full_${T}.log | cut -c1-200
