"""Writes tests/golden/r1_tangent_bits.json: sha256 digests of A and B of tests/r1_bits_cases.py as produced by the
device path of the library it is run with.  Run ONCE on a B200 with the round-1 library (commit 79635c4: node-major
element rows, brick_tangent_sym_kernel + assemble_A_kernel); later rounds must reproduce the digests bit for bit.

    gpurun -- 'python tests/golden/make_r1_bits.py gpurun_out/r1_tangent_bits.json'
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import xara_b200 as xb  # noqa: E402
from r1_bits_cases import CASES, run_case  # noqa: E402

out = {name: run_case(xb, name) for name in CASES}
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "r1_tangent_bits.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", dst, {k: v[2][:12] for k, v in out.items()})
