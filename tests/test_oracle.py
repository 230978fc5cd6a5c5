"""CPU: the oracle (oracle/xara_oracle.c) against the golden vectors generated from the
reference, and -- when oracle/_ref is present -- against the reference itself.

Tolerances: integer work (DOF ids, FE ids, sparse pattern) bit-exact; FP64 results within
1e-12 of the result's own scale (BASELINE.json north_star: "1e-12 relative FP64").
"""
import os

import numpy as np
import pytest

from golden_cases import CASES, DISPCONTROL_CASES, NSTEPS, RAYLEIGH_CASES, TRANSIENT_CASES, ele_nd, newmark_coeffs
from modelspec import (ELASTIC, J2_STEEL, MAT_ELASTIC, MAT_J2, ND_3D, ND_PLANE_STRAIN, ND_PLANE_STRESS, OracleBackend, RefBackend,
                       brick_block, brick_periodic_equaldof, disp_control, frame2d, frame2d_diaphragm_equaldof, frame3d, have_ref,
                       oracle_nd_path, oracle_uni_path, quad_plane, quad_plane_stress_pressure, ref_nd_path, soil_column_equaldof, tie)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12


def close(a, b, rtol=RTOL):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= rtol * scale


@pytest.mark.parametrize("key,kind,type_", [("j2_3d", MAT_J2, ND_3D), ("j2_pstrain", MAT_J2, ND_PLANE_STRAIN),
                                            ("elastic_3d", MAT_ELASTIC, ND_3D),
                                            ("elastic_pstrain", MAT_ELASTIC, ND_PLANE_STRAIN)])
def test_material_paths_vs_golden(key, kind, type_):
    g = np.load(os.path.join(GOLD, "material_paths.npz"))
    s, t = oracle_nd_path(kind, g[key + "_par"], type_, g[key + "_strain"], g[key + "_commit"])
    assert close(s, g[key + "_stress"])
    assert close(t, g[key + "_tangent"])
    if kind == MAT_J2:  # the path must actually yield, or the test says nothing about the return map
        el = g[key + "_tangent"][0]
        assert np.abs(g[key + "_tangent"] - el).max() > 1e-3 * np.abs(el).max()


@pytest.mark.parametrize("key,kind", [("steel02", 0), ("concrete02_core", 1), ("concrete02_cover", 1)])
def test_uniaxial_paths_vs_golden(key, kind):
    g = np.load(os.path.join(GOLD, "material_paths.npz"))
    s, t = oracle_uni_path(kind, g[key + "_par"], g[key + "_strain"], g[key + "_commit"])
    assert close(s, g[key + "_stress"], 1e-12)
    assert close(t, g[key + "_tangent"], 1e-11)       # Steel02's tangent goes through three pow() calls
    assert len(np.unique(np.round(g[key + "_tangent"], 3))) > (20 if kind == 0 else 3)     # the path really cycles


@pytest.mark.parametrize("name", list(CASES))
def test_model_vs_golden(name):
    mk, numberer, soe, _ = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = mk()
    O = OracleBackend(spec, numberer, soe)
    nd = ele_nd(spec)
    assert np.array_equal(O.ids(), g["ids"])                       # bit-exact DOF numbering
    ptr, idx = O.csr()
    assert np.array_equal(ptr, g["ptr"]) and np.array_equal(idx, g["idx"])   # bit-exact pattern
    assert np.array_equal(O.fe_ids(nd)[1], g["fe_ids"])
    for s in range(NSTEPS):
        O.set_trial_disp(g[f"u{s}"]); O.apply_load(0.25 * (s + 1))
        assert close(O.form_tangent(), g[f"A{s}"])
        assert close(O.form_unbalance(), g[f"B{s}"])
        for e in range(len(g[f"K{s}"])):
            assert close(O.ele_tangent(e, nd), g[f"K{s}"][e])
            assert close(O.ele_resid(e, nd), g[f"R{s}"][e])
        O.commit()


def test_scatter_map_is_where_addA_lands():
    """oracle scatter map == the positions its own addA loop writes (unit impulse per entry)."""
    spec = brick_block(2, 2, 1, distort=0.1)
    for soe in (0, 1):
        O = OracleBackend(spec, 1, soe)
        ptr, idx = O.csr()
        _, fe = O.fe_ids(24)
        for e in range(O.ne):
            m = O.scatter_map(e, 24)
            for i in range(24):
                for j in range(24):
                    r, c = fe[e, i], fe[e, j]
                    if r < 0 or c < 0:
                        assert m[i, j] == -1
                        continue
                    major, minor = (r, c) if soe == 1 else (c, r)
                    k = m[i, j]
                    assert ptr[major] <= k < ptr[major + 1] and idx[k] == minor


def test_empty_and_constrained_edge_cases():
    # every dof fixed: zero equations, empty pattern
    spec = brick_block(1, 1, 1)
    spec.fix = np.array([(t, d) for t in spec.node_tags for d in range(3)], np.int32)
    O = OracleBackend(spec, 0, 0)
    assert O.neq == 0 and O.nnz == 0
    # an isolated node (no element) keeps a diagonal-only row
    spec = brick_block(1, 1, 1)
    spec.node_tags = np.append(spec.node_tags, 100).astype(np.int32)
    spec.crd = np.vstack([spec.crd, [5.0, 5.0, 5.0]])
    O = OracleBackend(spec, 0, 1)
    ptr, idx = O.csr()
    ids = O.ids()
    for r in ids[-1]:
        assert ptr[r + 1] - ptr[r] == 1 and idx[ptr[r]] == r


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mat", [J2_STEEL, ELASTIC])
@pytest.mark.parametrize("numberer", [0, 1])
@pytest.mark.parametrize("soe", [0, 1])
def test_oracle_vs_live_reference(mat, numberer, soe):
    rng = np.random.default_rng(7)
    specs = [brick_block(3, 2, 2, mat=mat, distort=0.2, seed=11), quad_plane(5, 4, mat=mat, distort=0.2, seed=12)]
    if mat is J2_STEEL:
        specs.append(frame2d(2, 2, 2))
        specs.append(frame3d(1, 1, 2))
        specs.append(frame2d_diaphragm_equaldof(2, 2, 1))
    specs += [soil_column_equaldof(5, mat=mat), brick_periodic_equaldof(2, 2, 2, mat=mat)]     # `equalDOF`
    specs.append(quad_plane_stress_pressure(5, 3, 1 if mat is ELASTIC else 0, 1.5, mat=mat))      # PlaneStress (elastic), pressure
    for spec in specs:
        beam = spec.groups[0].kind in (2, 3)
        O, R = OracleBackend(spec, numberer, soe), RefBackend(spec, numberer, soe)
        assert O.neq == R.neq and O.nnz == R.nnz
        assert np.array_equal(O.ids(), R.ids())
        assert all(np.array_equal(a, b) for a, b in zip(O.csr(), R.csr()))
        for s in range(3):
            sc = (0.02, 0.02, 2e-4) if spec.groups[0].kind == 2 else ((0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4) if beam else 2e-3)
            u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * np.asarray(sc) * (s + 1); u[O.ids() < 0] = 0
            tie(spec, u)
            O.set_trial_disp(u); R.set_trial_disp(u)
            O.apply_load(0.3 * s); R.apply_load(0.3 * s)
            assert close(O.form_tangent(), R.form_tangent(), 1e-11 if beam else RTOL)
            assert close(O.form_unbalance(), R.form_unbalance(), 1e-11 if beam else RTOL)
            O.commit(); R.commit()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dim,loads", [(2, "uniform"), (2, "point"), (2, "both"), (3, "uniform"), (3, "point"), (3, "both"), (2, "partial"), (2, "all"), (3, "partial"), (3, "all")])
def test_beam_uniform_element_loads_vs_live_reference(dim, loads):
    """`eleLoad -beamPoint` (Beam2d/3dPointLoad: ForceBeamColumn2d.cpp:442-455, 1138-1181; ForceBeamColumn3d.cpp:457-475,
    1314-1373; points between and beyond the Lobatto sections) and `eleLoad -beamUniform` on force beams (ForceBeamColumn2d.cpp:407,1034 / ForceBeamColumn3d.cpp:419,1197): the section
    forces sp inside the element iteration and the fixed-end reactions p0 in the resisting force -- A, B and the
    element forces against the reference over a load history in which the load factor grows (gravity ramp), incl. steps
    with a zero displacement increment (the element must iterate all the same: numEleLoads > 0)"""
    from modelspec import with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(5)
    spec = frame2d(2, 2, 2) if dim == 2 else frame3d(1, 1, 2)
    # "partial" / "all": a trapezoidal load over part of every girder (Beam2d/3dPartialUniformLoad, ForceBeamColumn2d.cpp:426-443,
    # 1073-1137; ForceBeamColumn3d.cpp:432-456, 1224-1313), alone and on top of the other two kinds
    if loads not in ("point", "partial"): spec = with_beam_gravity(spec, seed=3)
    if loads not in ("uniform", "partial"): spec = with_beam_point_loads(spec, seed=2)
    if loads in ("partial", "all"):
        from modelspec import with_beam_partial_loads
        spec = with_beam_partial_loads(spec, seed=4)
    if loads == "both":       # a second uniform load (live on top of dead) on the loaded elements
        spec.beam_loads = spec.beam_loads + [(t, 0.4 * wy, 0.3 * wz, -0.5 * wa) for t, wy, wz, wa in spec.beam_loads[::2]]
    assert len(spec.beam_loads) + len(spec.beam_point_loads) + len(spec.beam_partial_loads) >= 4
    O, R = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0)
    sc = np.asarray((0.02, 0.02, 2e-4) if dim == 2 else (0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4))
    u = np.zeros((spec.nn, spec.ndf))
    for s_ in range(5):
        if s_ != 2:                     # step 2: the load factor moves, the displacements do not
            u = u + rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * 0.1; u[O.ids() < 0] = 0
        lam = 0.25 * (s_ + 1)
        for m in (O, R):
            m.apply_load(lam); m.set_trial_disp(u)
        assert close(O.form_tangent(), R.form_tangent(), 1e-11)
        assert close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        nd = 6 if dim == 2 else 12
        for e in range(O.ne):
            assert close(O.ele_resid(e, nd), R.ele_resid(e, nd), 1e-11)
        O.commit(); R.commit()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("numberer", [0, 1])
def test_band_and_profile_storage_vs_live_reference(numberer):
    """`system BandGeneral` / `system ProfileSPD`: the oracle's layout (numSubD / numSuperD, iDiagLoc) and its addA into the
    band / profile array against the reference's own BandGenLinSOE / ProfileSPDLinSOE sized from the same DOF graph and
    filled in FE_Element order -- layout bit-exact, A to rounding; equalDOF and constrained dofs included"""
    rng = np.random.default_rng(3)
    specs = [brick_block(3, 2, 2, distort=0.2, seed=11), quad_plane(5, 4, mat=J2_STEEL, distort=0.2, seed=12), frame2d(2, 2, 2),
             soil_column_equaldof(5), brick_periodic_equaldof(2, 2, 2)]
    for spec in specs:
        beam = spec.groups[0].kind in (2, 3)
        R = RefBackend(spec, numberer, 0)
        u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * np.asarray((0.02, 0.02, 2e-4) if beam else 2e-3); u[R.ids() < 0] = 0
        tie(spec, u)
        R.set_trial_disp(u); R.apply_load(0.4); R.form_tangent()
        for kind in (2, 3):
            O = OracleBackend(spec, numberer, kind)
            O.set_trial_disp(u); O.apply_load(0.4)
            lay, Ar = R.store_tangent(kind)
            if kind == 2:
                assert O.band() == lay
            else:
                assert np.array_equal(O.profile(), lay)
            Ao = O.form_tangent()
            assert len(Ao) == len(Ar) and close(Ao, Ar, 1e-11 if beam else RTOL)
            # the scatter map says where addA puts every element entry: scattering the element matrices through it
            # reproduces the array
            nd = {0: 24, 1: 8, 2: 6, 3: 12}[spec.groups[0].kind]
            acc = np.zeros(len(Ao))
            for e in range(O.ne):
                mp, K = O.scatter_map(e, nd).ravel(), O.ele_tangent(e, nd).ravel()
                np.add.at(acc, mp[mp >= 0], K[mp >= 0])
            assert close(acc, Ao, 1e-11 if beam else RTOL)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_revert_to_last_commit_vs_live_reference():
    """Domain::revertToLastCommit (Domain.cpp:1925): nodes and elements go back, the load factor of the last commit is
    applied again, then update() -- incl. J2PlaneStress's out-of-plane strain and the force beams' section state"""
    rng = np.random.default_rng(0)
    for spec, sc in ((brick_block(3, 3, 3, distort=0.1), 4e-3), (frame2d(2, 2, 2), np.array((0.006, 0.003, 6e-5))),
                     (quad_plane_stress_pressure(5, 4, 1, 1.5, mat=J2_STEEL), 3e-3)):
        O, R = OracleBackend(spec, 0, 1), RefBackend(spec, 0, 1)
        ids = O.ids()
        u1 = rng.normal(0, 1, (spec.nn, spec.ndf)) * sc; u1[ids < 0] = 0
        u2 = u1 + rng.normal(0, 1, (spec.nn, spec.ndf)) * sc; u2[ids < 0] = 0
        for m in (O, R):
            m.set_trial_disp(u1); m.apply_load(0.5); m.commit()
            m.set_trial_disp(u2); m.apply_load(0.9); m.revert()
        tol = 1e-11 if spec.groups[0].kind == 2 else RTOL
        assert close(O.form_unbalance(), R.form_unbalance(), tol)        # lambda = 0.5 again
        assert close(O.form_tangent(), R.form_tangent(), tol)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_revert_to_start_vs_live_reference():
    """Domain::revertToStart (Domain.cpp:1951): after a committed history every element kind is back at its initial
    state -- bitwise the initial A -- and the next step agrees with the reference"""
    rng = np.random.default_rng(2)
    cases = [(brick_block(3, 3, 3, distort=0.1), 4e-3), (quad_plane(5, 4, mat=J2_STEEL, lx=5.0, ly=4.0, distort=0.2), 3e-3),
             (quad_plane_stress_pressure(5, 4, 1, 1.5, mat=J2_STEEL), 3e-3), (quad_plane_stress_pressure(5, 4, 1, 1.5), 2e-2),
             (frame2d(2, 2, 2), np.array((0.006, 0.003, 6e-5))), (frame3d(1, 1, 2), np.array((0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4)))]
    for spec, sc in cases:
        tol = 1e-11 if spec.groups[0].kind in (2, 3) else RTOL
        O, R = OracleBackend(spec, 1, 1), RefBackend(spec, 1, 1)
        ids = O.ids()
        A0 = O.form_tangent().copy()
        for s in range(3):
            u = rng.normal(0, 1, (spec.nn, spec.ndf)) * sc * (s + 1); u[ids < 0] = 0
            for m in (O, R):
                m.set_trial_disp(u); m.apply_load(0.4 * (s + 1)); m.commit()
        for m in (O, R):
            m.revert_to_start()
        assert np.array_equal(O.form_tangent(), A0)
        assert close(O.form_tangent(), R.form_tangent(), tol) and np.abs(O.form_unbalance() - R.form_unbalance()).max() < 1e-12
        u = rng.normal(0, 1, (spec.nn, spec.ndf)) * sc; u[ids < 0] = 0
        for m in (O, R):
            m.set_trial_disp(u); m.apply_load(0.3)
        assert close(O.form_tangent(), R.form_tangent(), tol) and close(O.form_unbalance(), R.form_unbalance(), tol)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_material_paths_vs_live_reference():
    rng = np.random.default_rng(3)
    for kind, p in (J2_STEEL, ELASTIC):
        for type_ in (ND_3D, ND_PLANE_STRAIN):
            order = 6 if type_ == ND_3D else 3
            strains = np.cumsum(rng.normal(0, 1e-3, (120, order)), axis=0)
            commit = (rng.random(120) < 0.5).astype(np.int32)
            so, to = oracle_nd_path(kind, p, type_, strains, commit)
            sr, tr = ref_nd_path(kind, p, type_, strains, commit)
            assert close(so, sr) and close(to, tr)
    # the PlaneStress copies: ElasticIsotropicPlaneStress2D, and J2PlaneStress with its trial-to-trial out-of-plane strain
    for kind, p in (J2_STEEL, ELASTIC):
        strains = np.cumsum(rng.normal(0, 6e-4, (200, 3)), axis=0)
        commit = (rng.random(200) < 0.6).astype(np.int32)
        so, to = oracle_nd_path(kind, p, ND_PLANE_STRESS, strains, commit)
        sr, tr = ref_nd_path(kind, p, ND_PLANE_STRESS, strains, commit)
        assert close(so, sr) and close(to, tr)
        if kind == MAT_J2:
            assert len(np.unique(np.round(tr[:, 0, 0], 1))) > 50           # the path really goes plastic


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_steel01_elastic_paths_vs_live_reference():
    """Steel01 (Steel01.cpp:68-196, incl. the isotropic-hardening shifts a1..a4) and ElasticMaterial (bilinear Epos / Eneg)
    against the reference's own classes over cyclic strain paths with commits"""
    from modelspec import STEEL01_EX2B, UNI_STEEL01, UNI_ELASTIC, ref_uni_path
    rng = np.random.default_rng(4)
    from modelspec import UNI_CONCRETE01, UNI_ELASTICPP
    for kind, p in ((UNI_ELASTICPP, (2.9e4, 2.0e-3, -1.5e-3, 1.0e-4)), (UNI_ELASTICPP, (2.9e4, -1.0e-3, 1.0e-3, 0.0)), STEEL01_EX2B, (UNI_STEEL01, (60.0, 29000.0, 0.02, 0.05, 20.0, 0.04, 25.0)),
                    (UNI_ELASTIC, (3.0e4, 0.0, 3.0e4)), (UNI_ELASTIC, (3.0e4, 0.0, 1.0e4)),
                    (UNI_CONCRETE01, (-6.0, -0.004, -5.0, -0.014)), (UNI_CONCRETE01, (5.0, 0.002, 1.0, 0.006))):
        epsy = p[0] / p[1] if kind == UNI_STEEL01 else (1.5e-3 if kind == UNI_CONCRETE01 else 1e-3)
        if kind == UNI_ELASTICPP: epsy = 0.5e-3
        t = np.linspace(0, 14 * np.pi, 600)
        strains = 4.0 * epsy * (0.2 + t / t[-1]) * np.sin(t) + rng.normal(0, 0.05 * epsy, 600)
        strains[100:103] = strains[99]                      # zero increments (|dStrain| <= DBL_EPSILON: nothing moves)
        commit = (rng.random(600) < 0.7).astype(np.int32)
        so, to = oracle_uni_path(kind, p, strains, commit)
        sr, tr = ref_uni_path(kind, p, strains, commit)
        assert close(so, sr, 1e-14) and (close(to, tr, 1e-13) if kind == UNI_CONCRETE01 else np.array_equal(to, tr))   # (the reference is built with FMA contraction)
        if kind == UNI_CONCRETE01:
            assert (sr == 0.0).sum() > 50 and sr.min() < 0.9 * -abs(p[0]) and len(np.unique(np.round(tr, 2))) > 30   # cracks open, crushing, unloading slopes
        if kind == UNI_ELASTICPP:
            assert (tr == 0.0).sum() > 50 and (tr == p[0]).sum() > 50              # plastic plateaus and elastic unloading
        if kind == UNI_STEEL01:
            assert (tr == p[2] * p[1]).sum() > 50 and (tr == p[1]).sum() > 50      # the path yields and unloads


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("ndiv", [1, 3])
def test_aggregator_cantilever_vs_live_reference(ndiv):
    """BASELINE configs[0] as the script defines it: forceBeamColumn over `section Aggregator` (Elastic on P, Steel01 on
    Mz; SectionAggregator.cpp:316-500: diagonal tangent, flexibility 1/k) -- A, B, element forces against the reference
    over a pushed-and-released history with commits, a revert to the last commit and a reset"""
    from modelspec import cantilever2d_aggregator
    rng = np.random.default_rng(9)
    spec = cantilever2d_aggregator(ndiv=ndiv)
    O, R = OracleBackend(spec, 0, 0), RefBackend(spec, 0, 0)
    assert np.array_equal(O.ids(), R.ids())
    amp = np.array([10.0, 0.005, 0.04])          # the tip goes to 2.5 x its yield displacement
    yielded = False
    for s_, f in enumerate([0.1, 0.5, 1.0, 0.7, 0.2, -0.5, -1.0, 0.3]):
        h = np.linspace(0.0, 1.0, spec.nn)[:, None]
        u = f * amp * h ** 2 + rng.normal(0, 1.0, (spec.nn, 3)) * amp * 0.01; u[O.ids() < 0] = 0
        for m in (O, R):
            m.set_trial_disp(u); m.apply_load(0.1 * s_)
        Ao, Ar = O.form_tangent(), R.form_tangent()
        assert close(Ao, Ar, 1e-12) and close(O.form_unbalance(), R.form_unbalance(), 1e-12)
        for e in range(O.ne):
            assert close(O.ele_resid(e, 6), R.ele_resid(e, 6), 1e-12) and close(O.ele_tangent(e, 6), R.ele_tangent(e, 6), 1e-12)
        if s_ == 0: A0 = Ar.copy()
        yielded = yielded or abs(Ar[0] - A0[0]) > 0.5 * abs(A0[0])          # the lateral stiffness drops
        if s_ == 4:                       # a trial state thrown away
            O.revert(); R.revert()
            assert close(O.form_tangent(), R.form_tangent(), 1e-12) and close(O.form_unbalance(), R.form_unbalance(), 1e-12)
        else:
            O.commit(); R.commit()
    assert yielded
    O.revert_to_start(); R.revert_to_start()
    for m in (O, R):
        m.set_trial_disp(np.zeros((spec.nn, 3))); m.apply_load(0.0)
    assert close(O.form_tangent(), R.form_tangent(), 1e-12)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("bars", ["steel01", "elasticpp"])
@pytest.mark.parametrize("dim", [2, 3])
def test_steel01_elastic_fibres_vs_live_reference(dim, bars):
    """Steel01 bars and an Elastic (bilinear Epos / Eneg) cover inside FiberSection2d / FiberSection3d: the new uniaxial
    kinds as ordinary fibres, against the reference's classes over a sway history with commits"""
    from modelspec import steel01_elastic_frame
    rng = np.random.default_rng(8)
    spec = steel01_elastic_frame(dim, bars)          # "elasticpp": ElasticPPMaterial bars (the plastic strain moves at commitState)
    O, R = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0)
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    A0 = R.form_tangent().copy()
    # virgin state, every fibre at a strain of exactly zero: a fibre section asks ElasticMaterial::setTrial, whose tangent
    # there is Epos -- not getTangent()'s max(Epos, Eneg), which a section Aggregator gets (ElasticMaterial.cpp:146-182)
    assert close(O.form_tangent(), A0, 1e-12)
    for s_, a in enumerate([0.3, 0.8, 1.4, 2.0]):
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5
        u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        u += rng.normal(0, 1.0, u.shape) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))
        u[O.ids() < 0] = 0
        for m in (O, R):
            m.apply_load(0.25 * (s_ + 1)); m.set_trial_disp(u)
        assert close(O.form_tangent(), R.form_tangent(), 1e-11) and close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        O.commit(); R.commit()
    assert not close(R.form_tangent(), A0, 0.05)           # well past yield


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("pdelta", [0, 1])
def test_joint_offsets_vs_live_reference(pdelta, dim):
    """`geomTransf Linear | PDelta ... -jntOffset` (rigid end zones; the nodeIOffset / nodeJOffset terms of
    LinearCrdTransf2d.cpp and PDeltaCrdTransf2d.cpp): element length and orientation between the offset ends, basic
    deformations, tangent and resisting force pulled back to the nodes -- against the reference's classes over a sway
    history under gravity element loads, with commits and a revert"""
    from modelspec import with_joint_offsets, with_beam_gravity, with_pdelta
    rng = np.random.default_rng(21)
    mk = (lambda: with_beam_gravity(frame2d(2, 2, 2, gravity=-80.0), w=-0.08, seed=1)) if dim == 2 else \
         (lambda: with_beam_gravity(frame3d(1, 1, 2, gravity=-40.0), w=-0.06, seed=1))
    spec = with_joint_offsets(mk(), seed=3)
    if pdelta: spec = with_pdelta(spec)
    O, R, Rn = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0), RefBackend(with_pdelta(mk()) if pdelta else mk(), 1, 0)
    assert close(O.form_tangent(), R.form_tangent(), 1e-11)
    hc = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hc.max(); h = hc / H
    nd = 6 if dim == 2 else 12
    pattern = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))
    differs = False
    BT = 1e-11 if dim == 2 else 1e-10          # (the element state is the fixed point of an iteration converged to |dW| < 1e-12)
    for s_, a in enumerate([0.2, 0.5, 0.8, 1.1, 1.4] if dim == 2 else [0.2, 0.4, 0.6, 0.8, 1.0]):
        u = np.zeros((spec.nn, spec.ndf)); u[:, 0] = a * h ** 1.5; u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        if dim == 3: u[:, 1] = 0.5 * a * h ** 1.5; u[:, 3] = 0.7 * a * h ** 0.5 / H
        u += pattern * (a / 0.5); u[O.ids() < 0] = 0
        for m in (O, R, Rn):
            m.apply_load(0.2 * (s_ + 1)); m.set_trial_disp(u)
        Ar, Br = R.form_tangent(), R.form_unbalance()
        assert close(O.form_tangent(), Ar, BT) and close(O.form_unbalance(), Br, BT)
        for e in range(O.ne):
            assert close(O.ele_resid(e, nd), R.ele_resid(e, nd), BT) and close(O.ele_tangent(e, nd), R.ele_tangent(e, nd), BT)
        differs = differs or not close(Ar, Rn.form_tangent(), 1e-3)
        if s_ == 4:
            O.revert(); R.revert()
            assert close(O.form_tangent(), R.form_tangent(), BT) and close(O.form_unbalance(), R.form_unbalance(), BT)
        else:
            O.commit(); R.commit(); Rn.commit()
    assert differs


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dim", [2, 3])
def test_gravity_then_pushover_load_const_vs_live_reference(dim):
    """The usual RC-frame sequence: gravity (nodal loads and `eleLoad -beamUniform` / `-beamPoint`) ramped to its full
    value, `loadConst -time 0`, then a lateral pattern.  The frozen pattern keeps its factor for nodal and element loads
    alike (Domain::setLoadConstant, LoadPattern::applyLoad) while the new one follows the domain time."""
    from modelspec import with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(13)
    mk = (lambda: frame2d(2, 2, 2, lateral=0.0, gravity=-60.0)) if dim == 2 else (lambda: frame3d(1, 1, 2, lateral=(0.0, 0.0), gravity=-30.0))
    spec = with_beam_point_loads(with_beam_gravity(mk(), w=-0.08, seed=1), P=-2.0, seed=2)
    O, R = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0)
    ids = O.ids()
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    noise = (2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5)

    pattern = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * noise          # one irregular pattern, growing with the drift

    def step(a, lam):
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5; u[:, 1 if dim == 2 else 2] = -0.01 * h
        u += pattern * (a / 0.06); u[ids < 0] = 0
        for m in (O, R):
            m.apply_load(lam); m.set_trial_disp(u)
        assert close(O.form_tangent(), R.form_tangent(), 1e-11) and close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        for e in range(O.ne):
            assert close(O.ele_resid(e, nd), R.ele_resid(e, nd), 1e-11)
        B = R.form_unbalance().copy()
        O.commit(); R.commit()
        return B
    for s_ in range(3):
        Bg = step(0.02 * (s_ + 1), (s_ + 1) / 3.0)                 # gravity ramp
    O.load_const(); O.apply_load(0.0); R.load_const(0.0)
    top = [int(t) for t, c in zip(spec.node_tags, spec.crd) if c[1 if dim == 2 else 2] == H]
    for t in top:
        v = np.zeros(spec.ndf); v[0] = 12.0
        O.add_load(t, v); R.add_load(t, v)
    B0 = step(0.07, 0.0)                                           # time 0: the gravity loads are still there in full
    assert np.abs(B0).max() > 0.2 * np.abs(Bg).max()
    for s_ in range(3):
        step(0.3 * (s_ + 1), 0.3 * (s_ + 1))                       # the push


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("dim", [2, 3])
def test_beam_integration_rules_vs_live_reference(dim, kind):
    """forceBeamColumn -integration Legendre (1) | Radau (2) | NewtonCotes (3) | Trapezoidal (4): the reference builds its
    elements with that BeamIntegration class (quadrature/Frame/*BeamIntegration.cpp); the oracle takes the section locations
    and weights the class returns -- what the binding reads out of the element -- instead of its Lobatto tables.  Point and
    uniform element loads along (their section forces depend on the locations)."""
    from modelspec import with_beam_integration, with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(11)
    spec = frame2d(2, 2, 2, nip=4) if dim == 2 else frame3d(1, 1, 2, nip=5)
    spec = with_beam_integration(with_beam_point_loads(with_beam_gravity(spec, seed=3), seed=2), kind)
    lob = with_beam_point_loads(with_beam_gravity(frame2d(2, 2, 2, nip=4) if dim == 2 else frame3d(1, 1, 2, nip=5), seed=3), seed=2)
    O, R, Rl = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0), RefBackend(lob, 1, 0)
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    differs = False
    for s_, a in enumerate([0.2, 0.5, 0.8, 1.1]):
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5
        u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        u += rng.normal(0, 1.0, u.shape) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))
        u[O.ids() < 0] = 0
        for m in (O, R, Rl):
            m.apply_load(0.25 * (s_ + 1)); m.set_trial_disp(u)
        Ar, Br = R.form_tangent(), R.form_unbalance()
        assert close(O.form_tangent(), Ar, 1e-11) and close(O.form_unbalance(), Br, 1e-11)
        for e in range(O.ne):
            assert close(O.ele_resid(e, nd), R.ele_resid(e, nd), 1e-11)
        differs = differs or not close(Ar, Rl.form_tangent(), 1e-4)
        O.commit(); R.commit(); Rl.commit()
    assert differs                     # another rule, another tangent


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dim", [2, 3])
def test_pdelta_transformation_vs_live_reference(dim):
    """`geomTransf PDelta` under forceBeamColumn (PDeltaCrdTransf2d.cpp:349-745, PDeltaCrdTransf3d.cpp:200-249, 784-790,
    873-881): the geometric stiffness N/L and the leaning-column shear ul14 N/L against the reference's classes over a
    sway history under gravity (axial forces present), with commits and a revert to the last commit -- after which the 3D
    element keeps the relative displacements of its last update (ForceBeamColumn3d never refreshes the transformation
    in getResistingForce), the 2D one does not (ForceBeamColumn2d.cpp:402,526)"""
    from modelspec import with_pdelta, with_beam_gravity
    rng = np.random.default_rng(6)
    mk = (lambda: frame2d(2, 3, 2, gravity=-120.0)) if dim == 2 else (lambda: frame3d(1, 1, 2, gravity=-40.0))
    spec = with_pdelta(with_beam_gravity(mk(), seed=4))
    lin = with_beam_gravity(mk(), seed=4)
    O, R, Rl = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0), RefBackend(lin, 1, 0)
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    differs = False
    for s_, a in enumerate([0.3, 0.7, 1.1, 1.5, 1.9, 2.3]):          # inches of roof drift (the last step is thrown away)
        u = np.zeros((spec.nn, spec.ndf))
        if dim == 2:
            u[:, 0] = a * h ** 1.5; u[:, 1] = -0.01 * h; u[:, 2] = -1.5 * a * h ** 0.5 / H
            u += rng.normal(0, 1.0, u.shape) * (2e-3, 1e-3, 2e-5)
        else:
            u[:, 0] = a * h ** 1.5; u[:, 1] = 0.6 * a * h ** 1.5; u[:, 2] = -0.01 * h
            u[:, 3] = 0.9 * a * h ** 0.5 / H; u[:, 4] = -1.5 * a * h ** 0.5 / H
            u += rng.normal(0, 1.0, u.shape) * (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5)
        u[O.ids() < 0] = 0
        lam = 0.2 * (s_ + 1)
        for m in (O, R, Rl):
            m.apply_load(lam); m.set_trial_disp(u)
        Ao, Ar = O.form_tangent(), R.form_tangent()
        Bo, Br = O.form_unbalance(), R.form_unbalance()
        assert close(Ao, Ar, 1e-11) and close(Bo, Br, 1e-11)
        for e in range(O.ne):
            assert close(O.ele_resid(e, nd), R.ele_resid(e, nd), 1e-11) and close(O.ele_tangent(e, nd), R.ele_tangent(e, nd), 1e-11)
        differs = differs or (not close(Ar, Rl.form_tangent(), 1e-8) and not close(Br, Rl.form_unbalance(), 1e-8))   # (1000 x the parity tolerance)
        if s_ == 5:
            O.revert(); R.revert(); Rl.revert()
            assert close(O.form_tangent(), R.form_tangent(), 1e-11) and close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        else:
            O.commit(); R.commit(); Rl.commit()
    assert differs          # the P-Delta terms are really there: the Linear transformation gives another system


def drive_transient_vs_golden(model, g, name, check, is_dev=False):
    """replays the golden Newmark history (the reference's dU per iteration) through `model`"""
    (c1, c2, c3), (a1, a2, a3, a4) = newmark_coeffs(float(g["gamma"]), float(g["beta"]), float(g["dt"]))
    t = 0.0
    neq = len(g["B0_0"])
    if "rayleigh" in g:          # `rayleigh` command before the analysis: elements (Kc = current tangent) and nodes
        model.set_rayleigh(*[float(x) for x in g["rayleigh"]])
    for s in range(int(g["nsteps"])):
        t += float(g["dt"])
        model.set_transient(c1, c2, c3); model.newmark_predict(a1, a2, a3, a4); model.apply_load(t)
        model.incr_response(np.zeros(neq), 1.0, c2, c3)          # newStep ends with updateDomain(time, dT)
        for it in range(int(g["niter"])):
            B = model.form_unbalance(); A = model.form_tangent()
            check(A, g[f"A{s}_{it}"], B, g[f"B{s}_{it}"], np.abs(g[f"B{s}_0"]).max())
            model.incr_response(g[f"dU{s}_{it}"], 1.0, c2, c3)
        v, a = model.vel_accel()
        assert close(v, g[f"v{s}"], 1e-11) and close(a, g[f"a{s}"], 1e-11)
        model.commit()


@pytest.mark.parametrize("name", list(TRANSIENT_CASES) + list(RAYLEIGH_CASES))
def test_newmark_vs_golden(name):
    mk, mass_fn, *_ = {**TRANSIENT_CASES, **RAYLEIGH_CASES}[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = mk()
    O = OracleBackend(spec, 1, 1)
    O.set_mass(spec.node_tags, g["mass"])
    assert np.array_equal(O.ids(), g["ids"])

    def check(A, Ag, B, Bg, bscale):
        assert close(A, Ag, 1e-11)
        assert np.abs(B - Bg).max() <= 1e-11 * bscale          # B -> 0 as Newton converges: scale by the step's first residual

    drive_transient_vs_golden(O, g, name, check)


def test_ex2b_as_written_vs_golden():
    """BASELINE configs[0], tests/Ex2b.Canti2D.InelasticSection.Push.py as written (section Aggregator of Elastic + Steel01,
    gravity under LoadControl, loadConst -time 0, pushover under DisplacementControl; golden: the reference's own classes
    with `system BandGeneral`, tests/golden/make_golden.py ex2b_as_written) against the same sequence on the oracle"""
    from modelspec import cantilever2d_aggregator, ex2b_drive
    g = np.load(os.path.join(GOLD, "ex2b_as_written.npz"))
    spec = cantilever2d_aggregator(ndiv=1, H=0.0, V=-2000.0)
    O = OracleBackend(spec, 0, 0)
    ids = O.ids(); ptr, idx = O.csr(); neq = O.neq

    def solve(A, b):
        M = np.zeros((neq, neq))
        for c in range(neq): M[idx[ptr[c]:ptr[c + 1]], c] = A[ptr[c]:ptr[c + 1]]
        return np.linalg.solve(M, b)

    grav_u, lam, u = ex2b_drive(O, solve, ids, False)
    assert close(grav_u, g["grav_u"], 1e-10)
    assert close(lam, g["push_lam"], 1e-9) and close(u, g["push_u"], 1e-9)
    assert lam[-1] < 0.5 * lam[8] * 50 / 9           # the section yielded: the push curve flattens


@pytest.mark.parametrize("name", list(DISPCONTROL_CASES))
def test_displacement_control_vs_golden(name):
    """`integrator DisplacementControl` + Newton run by the reference's own classes (golden) against the
    same algorithm driving the oracle's update / formUnbalance / formTangent: identical iteration
    counts on every step, the same load-factor history and final displacements."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    mk, numberer, soe, node, dof, incr, nsteps, tol, max_iter = DISPCONTROL_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = mk()
    O = OracleBackend(spec, numberer, soe); O._u = np.zeros((spec.nn, spec.ndf))
    ids = O.ids()
    assert np.array_equal(ids, g["ids"])
    ptr, idx = O.csr(); neq = O.neq

    def solve(A, b):   # soe 0 is column-compressed: the stored pattern is that of A^T
        M = sp.csr_matrix((A, idx, ptr), shape=(neq, neq))
        return spla.spsolve((M.T if soe == 0 else M).tocsc(), b)

    ctrl = ids[list(spec.node_tags).index(int(g["node"])), dof]
    hist, lam = disp_control(O, solve, ctrl, incr, nsteps, tol, max_iter, False)
    assert [len(h) for h in hist] == g["iters"].tolist()
    assert close(lam, g["lam"], 1e-9)
    assert close(O._u, g["u"], 1e-9)
    for s, h in enumerate(hist):
        assert np.allclose(h[:-1], g["norms"][s, :len(h) - 1], rtol=1e-5, atol=1e-12)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("numberer,soe", [(0, 0), (1, 1)])
def test_mixed_ndf_soil_frame_vs_live_reference(numberer, soe):
    """BASELINE configs[4] as a real mixed-ndf Domain: FourNodeQuad soil on 2-dof nodes (FourNodeQuad.cpp:133-139 accepts
    nothing else) carrying a forceBeamColumn frame on 3-dof nodes, the column bases tied to the soil surface with
    `equalDOF`.  DOF ids, FE_Element ids, pattern bit-exact; A, B to rounding over a load history."""
    from modelspec import soil_frame_2d
    rng = np.random.default_rng(11)
    spec = soil_frame_2d()
    O, R = OracleBackend(spec, numberer, soe), RefBackend(spec, numberer, soe)
    assert O.neq == R.neq and O.nnz == R.nnz
    ids = O.ids()
    assert np.array_equal(ids, R.ids())
    assert (ids[[i for i, t in enumerate(spec.node_tags) if int(t) in spec.node_ndf], 2] == -1).all()   # no third dof on soil nodes
    assert all(np.array_equal(a, b) for a, b in zip(O.csr(), R.csr()))
    to, io = O.fe_ids(8); tr, ir = R.fe_ids(8)
    assert len(to) == len(tr) and np.array_equal(io, ir)          # (FE_Element tags count from 0, element tags from 1)
    sc = np.array((0.02, 0.02, 2e-4))
    for s in range(3):
        u = rng.normal(0, 1.0, (spec.nn, 3)) * sc * 0.3 * (s + 1); u[ids < 0] = 0
        tie(spec, u)
        O.set_trial_disp(u); R.set_trial_disp(u)
        O.apply_load(0.3 * (s + 1)); R.apply_load(0.3 * (s + 1))
        assert close(O.form_tangent(), R.form_tangent(), 1e-11)
        assert close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        O.commit(); R.commit()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("loads", [0, 1])
def test_corotational_transformation_vs_live_reference(loads):
    """`geomTransf Corotational` on 2D force beams (CorotCrdTransf2d.cpp: basic deformations with the rigid-body rotation of
    the chord taken out, exact basic-to-local force transformation, geometric stiffness): a sway history with LARGE
    displacements (storey drifts of several per cent plus rotations), commits and a revert, against the reference's
    classes -- assembled A and B, every element's tangent and resisting force"""
    from modelspec import with_corot, with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(33)
    def mk():
        sp = frame2d(2, 2, 2, gravity=-80.0)
        if loads: sp = with_beam_point_loads(with_beam_gravity(sp, w=-0.08, seed=1), P=-2.0, seed=2)
        return sp
    spec = with_corot(mk())
    O, R, Rn = OracleBackend(spec, 1, 0), RefBackend(spec, 1, 0), RefBackend(mk(), 1, 0)
    assert close(O.form_tangent(), R.form_tangent(), 1e-11)
    H = spec.crd[:, 1].max(); h = spec.crd[:, 1] / H
    pattern = rng.normal(0, 1.0, (spec.nn, 3)) * (2e-3, 1e-3, 2e-5)
    differs = False
    for s_, a in enumerate([0.5, 1.5, 3.0, 5.0, 7.0, 5.5]):
        u = np.zeros((spec.nn, 3)); u[:, 0] = a * h ** 1.5; u[:, 2] = -1.5 * a * h ** 0.5 / H
        u += pattern * (a / 0.5); u[O.ids() < 0] = 0
        for m in (O, R, Rn):
            m.apply_load(0.2 * (s_ + 1)); m.set_trial_disp(u)
        Ar, Br = R.form_tangent(), R.form_unbalance()
        assert close(O.form_tangent(), Ar, 1e-11) and close(O.form_unbalance(), Br, 1e-11)
        for e in range(O.ne):
            assert close(O.ele_resid(e, 6), R.ele_resid(e, 6), 1e-11) and close(O.ele_tangent(e, 6), R.ele_tangent(e, 6), 1e-11)
        differs = differs or not close(Ar, Rn.form_tangent(), 1e-3)
        if s_ == 4:
            O.revert(); R.revert()
            assert close(O.form_tangent(), R.form_tangent(), 1e-11) and close(O.form_unbalance(), R.form_unbalance(), 1e-11)
        else:
            O.commit(); R.commit(); Rn.commit()
    assert differs


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("numberer", [0, 1])
def test_transformation_handler_numbers_and_assembles_as_plain(numberer):
    """`constraints Transformation` with homogeneous SPs and `equalDOF` constraints (identity constraint matrices): the
    reference's TransformationConstraintHandler gives the SAME equation numbers and SparseGenCol pattern as its PlainHandler
    (same DOF_Group / FE_Element tags and DOF-group graph; a constrained node's dofs take the retained node's equations
    either way) and, state for state, bitwise the same A and B for the continuum elements.  But its enforceSPs() updates
    every element next to a constrained node once more at each applyLoad (TransformationConstraintHandler.cpp:462-483): after
    a commit that turns the consistent tangent of a yielded J2 point into the elastic one and makes a force beam iterate
    again, so the reference's OWN Newton histories differ between its two handlers once the model yields (asserted at the
    end).  The device path follows PlainHandler by default and repeats the handler's second update on the same elements when the
    binding finds a Transformation handler (`constraints_transformation`, tests/test_gpu_parity.py `...@T`)."""
    rng = np.random.default_rng(17)
    specs = [soil_column_equaldof(6, mat=J2_STEEL, distort=0.1), brick_periodic_equaldof(3, 2, 2), brick_block(3, 3, 3, mat=J2_STEEL, distort=0.2, seed=2),
             quad_plane(6, 4, mat=J2_STEEL, distort=0.2, seed=3)]
    for spec in specs:
        P, T = RefBackend(spec, numberer, 0, handler=0), RefBackend(spec, numberer, 0, handler=1)
        assert P.neq == T.neq and P.nnz == T.nnz
        (pp, pi), (tp, ti) = P.csr(), T.csr()
        assert np.array_equal(pp, tp) and np.array_equal(pi, ti)
        _, fp = P.fe_ids(); _, ft = T.fe_ids()
        for a, b in zip(fp, ft):                    # the same equations per FE_Element (a TransformationFE lists them in its own order)
            assert sorted(set(a.tolist()) - {-9, -1}) == sorted(set(b.tolist()) - {-9, -1})
        for s_ in range(3):
            u = rng.normal(0, 2e-3 * (s_ + 1), (spec.nn, spec.ndf)); u[P.ids() < 0] = 0; tie(spec, u)
            for m in (P, T):
                m.set_trial_disp(u); m.apply_load(0.3 * (s_ + 1))
            assert np.array_equal(P.form_tangent(), T.form_tangent()) and np.array_equal(P.form_unbalance(), T.form_unbalance())
            P.commit(); T.commit()
    # force beams: equal numbering and pattern, values to the element tolerance only
    spec = frame2d(2, 2, 2)
    P, T = RefBackend(spec, numberer, 0, handler=0), RefBackend(spec, numberer, 0, handler=1)
    assert np.array_equal(P.ids(), T.ids()) and all(np.array_equal(a, b) for a, b in zip(P.csr(), T.csr()))
    u = rng.normal(0, 1.0, (spec.nn, 3)) * (0.02, 0.02, 2e-4); u[P.ids() < 0] = 0
    for m in (P, T):
        m.set_trial_disp(u); m.apply_load(0.4)
    assert close(P.form_tangent(), T.form_tangent(), 1e-5)
    # the same load-controlled Newton analysis under both handlers: identical while elastic, different once the soil yields
    def mk():
        sp = soil_column_equaldof(12, mat=J2_STEEL, distort=0.1)
        sp.loads = np.array([[1 + 2 * 12, 22.0 * 8, -3.0 * 8], [1 + 2 * 6, 10.0 * 8, 0.0]]); return sp
    hist = []
    for h in (0, 1):
        C = RefBackend(mk(), numberer, 0, dlambda=1.0 / 8, test=0, tol=1e-7, max_iter=25, handler=h)
        rc, iters, norms = C.analyze_static(8)
        assert rc == 0
        hist.append(norms)
    assert np.allclose(hist[0][:4], hist[1][:4], rtol=1e-9, atol=1e-14)
    assert not np.allclose(hist[0][5:, 0], hist[1][5:, 0], rtol=1e-5)
