#!/bin/bash
mkdir -p gpurun_out

run() { # label, env...
  python - "$@" <<'PY'
import json, os, subprocess, sys
label = sys.argv[1]; env = dict(os.environ); 
for kv in sys.argv[2:]:
    k, v = kv.split("=", 1); env[k] = v
r = subprocess.run([sys.executable, "bench.py", "--workload", "frame3d", "--n", "20", "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "1"], env=env, capture_output=True, text=True)
try:
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(label, 'ms/step', round(d['ms_per_step'], 3), 'update', round(d['kernel_ms']['update'], 3), 'roof', round(d['roofline']['frac'], 3), d['roofline']['kernel'])
except Exception as e:
    print(label, 'failed', e, r.stderr[-500:])
PY
}

for occ in 3 4 5; do run occ$occ XARA_B200_LIB=$PWD/scratch/lib_fbc_occ$occ.so; done
run default6
timeout 900 python -m pytest tests -m gpu -x -q -k "frame or golden or newmark or displacement" 2>&1 | tail -5
