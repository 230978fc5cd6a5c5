#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scratch/try_variants.sh 2>&1 | tee gpurun_out/asm_variants.log
