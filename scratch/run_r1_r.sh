#!/bin/bash
# secondary workloads (frame3d, frame) + ncu launch list of the headline bench + full captures
mkdir -p gpurun_out
for w in "frame3d 20" "frame 200"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r1_$1.err > gpurun_out/bench_r1_$1.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r1_$1.json'))
    print('$1', 'ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['kernel_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'elements', d['config']['elements'])
except Exception as e:
    print('$1 failed', e)
PY
  tail -2 gpurun_out/bench_r1_$1.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_q.log 2>&1
tail -2 gpurun_out/ncu_launch_q.log | cut -c1-300
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'brick_tangent_sym|brick_update|assemble_A' -c 3 -o gpurun_out/prof_r1_q python bench.py --n 96 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_q.log 2>&1
tail -2 gpurun_out/ncu_full_q.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fbc3d_update' -s 3 -c 1 -o gpurun_out/prof_r1_fbc3d python bench.py --workload frame3d --n 10 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_fbc3d.log 2>&1
tail -2 gpurun_out/ncu_full_fbc3d.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
