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

from golden_cases import CASES, NSTEPS, control_node, ele_nd  # noqa: E402
from modelspec import (CONCRETE02_CORE, CONCRETE02_COVER, ELASTIC, J2_STEEL, ND_3D, ND_PLANE_STRAIN, STEEL02,  # noqa: E402
                       RefBackend, ref_nd_path, ref_uni_path, tie)


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
    for name, (kind, p) in (("steel02", STEEL02), ("concrete02_core", CONCRETE02_CORE), ("concrete02_cover", CONCRETE02_COVER)):
        n = 300
        strains = np.cumsum(rng.normal(0, 6e-4, n)) - (0.002 if kind else 0.0)
        commit = (rng.random(n) < 0.6).astype(np.int32)
        s, t = ref_uni_path(kind, p, strains, commit)
        out[name + "_par"] = np.array(p); out[name + "_strain"] = strains; out[name + "_commit"] = commit
        out[name + "_stress"] = s; out[name + "_tangent"] = t
    np.savez_compressed(os.path.join(HERE, "material_paths.npz"), **out)


def model_case(name, spec, numberer, soe, scale, nsteps=NSTEPS):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    R = RefBackend(spec, numberer, soe)
    ids = R.ids()
    ptr, idx = R.csr()
    out = dict(ids=ids, ptr=ptr, idx=idx, numberer=numberer, soe=soe)
    nd = ele_nd(spec)
    _, fe = R.fe_ids(nd)
    out["fe_ids"] = fe
    for s in range(nsteps):
        u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * np.asarray(scale) * (s + 1); u[ids < 0] = 0
        tie(spec, u)
        R.set_trial_disp(u); R.apply_load(0.25 * (s + 1))
        out[f"u{s}"] = u
        out[f"A{s}"] = R.form_tangent(); out[f"B{s}"] = R.form_unbalance()
        out[f"K{s}"] = np.stack([R.ele_tangent(e, nd) for e in range(min(R.ne, 6))])
        out[f"R{s}"] = np.stack([R.ele_resid(e, nd) for e in range(min(R.ne, 6))])
        R.commit()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def transient_case(name, spec, mass, gamma, beta, dt, nsteps=3, niter=3, rayleigh=None):
    """Newmark steps driven through the reference's own integrator: newStep, then `niter` Newton
    iterations (formUnbalance, formTangent, solve, update) per step; everything recorded."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    R = RefBackend(spec, 1, 1, defer_setup=True)
    R.set_mass(spec.node_tags, mass)
    if rayleigh is not None:
        R.set_rayleigh(*rayleigh)
    R.setup_transient(1, 1, gamma, beta)
    ptr, idx = R.csr(); neq = R.neq
    out = dict(mass=mass, gamma=gamma, beta=beta, dt=dt, nsteps=nsteps, niter=niter, ids=R.ids(), ptr=ptr, idx=idx)
    if rayleigh is not None:
        out["rayleigh"] = np.array(rayleigh)
    for s in range(nsteps):
        assert R.new_step(dt) == 0
        for it in range(niter):
            B = R.form_unbalance(); A = R.form_tangent()
            dU = spla.spsolve(sp.csr_matrix((A, idx, ptr), shape=(neq, neq)).tocsc(), B)
            out[f"A{s}_{it}"] = A; out[f"B{s}_{it}"] = B; out[f"dU{s}_{it}"] = dU
            R.transient_update(dU)
        v, a = R.vel_accel()
        out[f"v{s}"] = v; out[f"a{s}"] = a; out[f"u{s}"] = R.get_trial_disp()
        R.commit()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def dispcontrol_case(name, spec, numberer, soe, node, dof, incr, nsteps, tol, max_iter):
    """the reference's StaticAnalysis loop with its own DisplacementControl, NewtonRaphson and
    CTestNormDispIncr (ref_analyze_static_lam): iteration counts, test norms, load factors, final state"""
    R = RefBackend(spec, defer_setup=True)
    node = control_node(spec, node)
    R.setup_dispcontrol(numberer, soe, node, dof, incr, test=0, tol=tol, max_iter=max_iter)
    rc, iters, norms, lam = R.analyze_static_lam(nsteps)
    assert rc == 0, rc
    # an iteration count must not hinge on the last bits of a norm: every deciding norm is >= 1.3x away
    # from tol (two implementations that agree to 1e-10 cannot then disagree on a count)
    for s in range(nsteps):
        last = norms[s, iters[s] - 1]
        assert last * 1.3 <= tol and (iters[s] == 1 or norms[s, iters[s] - 2] >= 1.3 * tol), (name, s, norms[s])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), iters=iters, norms=norms, lam=lam, u=R.get_trial_disp(),
                        ids=R.ids(), node=node, dof=dof, incr=incr, tol=tol, max_iter=max_iter, numberer=numberer, soe=soe)
    print(name, "iters", iters.tolist(), "lambda_end", lam[-1])


if __name__ == "__main__":
    # python tests/golden/make_golden.py [substring]: only the cases whose name contains it
    from golden_cases import DISPCONTROL_CASES, RAYLEIGH_CASES, TRANSIENT_CASES
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, (mk, mass_fn, gamma, beta, dt, ray) in RAYLEIGH_CASES.items():
        if only in name:
            spec = mk()
            transient_case(name, spec, mass_fn(spec), gamma, beta, dt, rayleigh=ray)
    for name, (mk, *args) in DISPCONTROL_CASES.items():
        if only in name:
            dispcontrol_case(name, mk(), *args)
    for name, (mk, mass_fn, gamma, beta, dt) in TRANSIENT_CASES.items():
        if only in name:
            spec = mk()
            transient_case(name, spec, mass_fn(spec), gamma, beta, dt)
    if only in "ex2b_as_written":
        from modelspec import ex2b_as_written
        g = ex2b_as_written(glue=False)
        np.savez_compressed(os.path.join(HERE, "ex2b_as_written.npz"), **g)
        print("ex2b_as_written gravity iters", g["grav_iters"].tolist(), "push iters", g["push_iters"].tolist(), "lambda_end", g["push_lam"][-1])
    if only in "material_paths":
        material_paths()
    for name, (mk, numberer, soe, scale) in CASES.items():
        if only in name:
            model_case(name, mk(), numberer, soe, scale)
    print("golden fixtures written to", HERE)
