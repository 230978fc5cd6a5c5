#!/bin/bash
# round-2 evidence: N=1 bench line, ncu launch list of the same command at n=160, ncu --set full of the four kernels at n=96
cd /root/repo
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_n160.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'brick_update|assemble_B|brick_tangent_rec|assemble_A' --launch-skip 13 --launch-count 4 -f \
  -o gpurun_out/r2_full_n96 python bench.py --n 96 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --opt ranged_tangent=0 > gpurun_out/r2_full.log 2>&1
tail -2 gpurun_out/r2_full.log | cut -c1-200
ls -la gpurun_out/r2_*
