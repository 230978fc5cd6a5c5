"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libref_harness.so,
built from /root/reference by oracle/ref_build.mk).  Run here, in the container that has
/root/reference; the fixtures travel to the GPU box where the reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import zlib  # noqa: E402

from golden_cases import CASES, NSTEPS  # noqa: E402
from modelspec import ELASTIC, J2_STEEL, ND_3D, ND_PLANE_STRAIN, RefBackend, ref_nd_path  # noqa: E402


def material_paths():
    rng = np.random.default_rng(20261017)
    out = {}
    for name, (kind, p) in (("j2", J2_STEEL), ("elastic", ELASTIC)):
        for tname, type_ in (("3d", ND_3D), ("pstrain", ND_PLANE_STRAIN)):
            order = 6 if type_ == ND_3D else 3
            n = 80
            strains = np.cumsum(rng.normal(0, 7e-4, (n, order)), axis=0)
            commit = (rng.random(n) < 0.7).astype(np.int32)
            s, t = ref_nd_path(kind, p, type_, strains, commit)
            key = f"{name}_{tname}"
            out[key + "_par"] = np.array(p); out[key + "_strain"] = strains; out[key + "_commit"] = commit
            out[key + "_stress"] = s; out[key + "_tangent"] = t
    np.savez_compressed(os.path.join(HERE, "material_paths.npz"), **out)


def model_case(name, spec, numberer, soe, scale, nsteps=NSTEPS):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    R = RefBackend(spec, numberer, soe)
    ids = R.ids()
    ptr, idx = R.csr()
    out = dict(ids=ids, ptr=ptr, idx=idx, numberer=numberer, soe=soe)
    nd = 24 if spec.ndm == 3 else 8
    _, fe = R.fe_ids(nd)
    out["fe_ids"] = fe
    for s in range(nsteps):
        u = rng.normal(0, scale * (s + 1), (spec.nn, spec.ndf)); u[ids < 0] = 0
        R.set_trial_disp(u); R.apply_load(0.25 * (s + 1))
        out[f"u{s}"] = u
        out[f"A{s}"] = R.form_tangent(); out[f"B{s}"] = R.form_unbalance()
        out[f"K{s}"] = np.stack([R.ele_tangent(e, nd) for e in range(min(R.ne, 6))])
        out[f"R{s}"] = np.stack([R.ele_resid(e, nd) for e in range(min(R.ne, 6))])
        R.commit()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    material_paths()
    for name, (mk, numberer, soe, scale) in CASES.items():
        model_case(name, mk(), numberer, soe, scale)
    print("golden fixtures written to", HERE)
