"""Cases whose assembled A (and B) are pinned BIT FOR BIT to what the round-1 device path produced on a B200
(tests/golden/r1_tangent_bits.json, written by tests/golden/make_r1_bits.py with the library of commit 79635c4).
The round-2 formTangent (symmetric element records + gathered assembly) must reproduce them exactly: same block
values, same FE_Element order of additions.  Static analysis only (the transient passes add in another order)."""
import hashlib

import numpy as np

from modelspec import ELASTIC, J2_STEEL, brick_block, brick_periodic_equaldof, tie


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def _shuffled(spec, seed=0):
    g = spec.groups[0]
    p = np.random.default_rng(seed).permutation(len(g.tags))
    g.tags, g.conn, g.mat, g.par = g.tags[p], g.conn[p], g.mat[p], g.par[p]
    return spec


CASES = {
    # name: (spec maker, numberer, soe, displacement scale)
    "brick_j2_plain_csc": (lambda: brick_block(5, 4, 3, mat=J2_STEEL, distort=0.25, seed=3, body=(0.01, 0.0, -0.02)), 0, 0, 1.5e-3),
    "brick_j2_rcm_csr": (lambda: brick_block(5, 4, 3, mat=J2_STEEL, distort=0.25, seed=3, body=(0.01, 0.0, -0.02)), 1, 1, 1.5e-3),
    "brick_elastic_rcm_csc": (lambda: brick_block(4, 6, 3, mat=ELASTIC, distort=0.2, seed=5), 1, 0, 1.5e-3),
    "brick_j2_shuffled_rcm_csc": (lambda: _shuffled(brick_block(14, 11, 9, mat=J2_STEEL, distort=0.2, seed=3)), 1, 0, 2.0e-3),
    "brick_j2_equaldof_rcm_csr": (lambda: brick_periodic_equaldof(5, 4, 3, seed=32), 1, 1, 1.5e-3),
    "brick_j2_equaldof_plain_csc": (lambda: brick_periodic_equaldof(3, 5, 4, dofs=(0, 1, 2), seed=33), 0, 0, 1.5e-3),
    # large enough for the ranged (pipelined) formTangent with a host destination
    "brick_j2_72k_plain_csr": (lambda: brick_block(48, 40, 38, mat=J2_STEEL, distort=0.15, seed=7), 0, 1, 2.5e-3),
}
NSTEPS = 3


def run_case(xb, name, options=None):
    """-> list of sha256 digests: A, B of the untouched model, then of every step of the load history"""
    mk, numberer, soe, sc = CASES[name]
    spec = mk()
    D = (xb.DeviceModel.from_spec(spec, numberer, soe, options=options) if options else xb.DeviceModel.from_spec(spec, numberer, soe)).to_device(0)
    ids = D.ids()
    rng = np.random.default_rng(1234)
    out = [_digest(D.form_tangent()), _digest(D.form_unbalance())]
    for s in range(NSTEPS):
        u = rng.normal(0, sc * (s + 1), (spec.nn, 3)); u[ids < 0] = 0
        tie(spec, u)
        D.set_trial_disp(u); D.update(); D.apply_load(0.3 * (s + 1))
        out += [_digest(D.form_tangent()), _digest(D.form_unbalance())]
        if s % 2 == 0:
            D.commit()
    return out
