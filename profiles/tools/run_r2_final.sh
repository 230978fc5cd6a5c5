#!/bin/bash
# the driver's command, then ncu --set full of the beam update kernel (frame3d workload)
cd /root/repo
( time python bench.py ) > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
tail -4 gpurun_out/r2f_bench_default.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'fbc3d_update_sec' --launch-skip 4 --launch-count 1 -f \
  -o gpurun_out/r2f_full_frame3d python bench.py --workload frame3d --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2f_full_frame3d.log 2>&1
tail -2 gpurun_out/r2f_full_frame3d.log | cut -c1-200
python profiles/tools/ncu_summary.py gpurun_out/r2f_full_frame3d.ncu-rep > gpurun_out/r2f_ncu_frame3d.txt 2>&1; cat gpurun_out/r2f_ncu_frame3d.txt
