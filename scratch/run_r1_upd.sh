#!/bin/bash
mkdir -p gpurun_out
bash scratch/try_variants.sh 2>&1 | tee gpurun_out/upd_variants.log
XARA_B200_LIB=$PWD/scratch/lib_upd_s1_o4.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'brick_tangent_sym|assemble_A' -c 2 -o gpurun_out/prof_r1_sym2 python bench.py --n 96 --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_sym2.log 2>&1
tail -2 gpurun_out/ncu_sym2.log
