#!/bin/bash
cd /root/repo
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'quad_|assemble_A|assemble_B' -s 10 -c 5 -o gpurun_out/prof_quad python bench.py --workload quad --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_quad.log 2>&1
tail -2 gpurun_out/ncu_quad.log | cut -c1-160
