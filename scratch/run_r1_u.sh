#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # label, workload, n, env...
  python - "$@" <<'PY'
import json, os, subprocess, sys
label, wl, n = sys.argv[1:4]; env = dict(os.environ)
for kv in sys.argv[4:]:
    k, v = kv.split("=", 1); env[k] = v
r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--n", n, "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "2"], env=env, capture_output=True, text=True)
try:
    d = json.loads(r.stdout.strip().splitlines()[-1])
    open(f"gpurun_out/bench_r1_{label}.json", "w").write(json.dumps(d) + "\n")
    print(label, 'ms/step', round(d['ms_per_step'], 3), 'update', round(d['kernel_ms']['update'], 3), 'roof', round(d['roofline']['frac'], 3), d['roofline']['kernel'], 'e2e', round(d['e2e']['ms_per_step'], 2))
except Exception as e:
    print(label, 'failed', e, r.stderr[-500:])
PY
}
run frame_element frame 200 XB_BEAM=element
run frame frame 200
run frame3d frame3d 20
