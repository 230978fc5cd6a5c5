#!/bin/bash
cd /root/repo
XB_TILE=2368 XB_FUSED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'fused' -s 3 -c 1 -o gpurun_out/prof_fused2 python bench.py --n 96 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log | cut -c1-200
