#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fbc3d_update_sec' -s 4 -c 1 -o gpurun_out/prof_r1_fbc3d_sec python bench.py --workload frame3d --n 20 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_fbc3d_sec.log 2>&1
tail -2 gpurun_out/ncu_full_fbc3d_sec.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fbc2d_update' -s 4 -c 1 -o gpurun_out/prof_r1_fbc2d python bench.py --workload frame --n 200 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_fbc2d.log 2>&1
tail -2 gpurun_out/ncu_full_fbc2d.log | cut -c1-200
