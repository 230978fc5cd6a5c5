#!/bin/bash
for f in scratch/lib_*.so; do
  v=$(basename $f .so)
  XARA_B200_LIB=$PWD/$f python bench.py --n 128 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null > /tmp/out.json
  python -c "
import json; d=json.load(open('/tmp/out.json')); print('$v', 'asmA', round(d['kernel_ms']['assemble_A'],3), 'tan', round(d['kernel_ms']['element_tangent'],3), 'upd', round(d['kernel_ms']['update'],3), 'res', round(d['kernel_ms']['element_resid'],3))"
done
