#!/bin/bash
# ncu --set full of the formTangent kernels (one launch each) at n^3 elements; $1 = tag, $2 = n, $3 = kernel regex
cd /root/repo
TAG=${1:-r2}; N=${2:-96}; RX=${3:-assemble_A_rec|brick_tangent_rec}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$RX" --launch-skip 6 --launch-count 2 -f -o gpurun_out/prof_$TAG \
  python bench.py --n $N --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
python profiles/tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_summary_$TAG.txt 2>&1
cat gpurun_out/ncu_summary_$TAG.txt
