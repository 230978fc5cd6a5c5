"""GPU: the CUDA path (through the C-ABI, xara_b200.DeviceModel) against the oracle and the
golden vectors generated from the reference.

Bar (BASELINE.json north_star): DOF numbering and scatter maps bit-exact (tests/test_host_setup.py),
element forces / tangents and the assembled A, B within 1e-12 relative, identical Newton
iteration counts.  "Relative" is to the largest entry of the object compared (an element
matrix, A, B): entries that cancel to ~0 cannot agree to 1e-12 of themselves between two
compilers of the reference either (FMA contraction).
"""
import os

import numpy as np
import pytest

import xara_b200 as xb
from golden_cases import CASES, NSTEPS, ele_nd
from r1_bits_cases import CASES as R1_CASES, run_case as r1_run_case
from modelspec import (ELASTIC, J2_STEEL, OracleBackend, brick_block, brick_periodic_equaldof, frame2d, frame2d_diaphragm_equaldof, frame3d,
                       have_glue, have_metis, have_ref, metis_partition, quad_plane, quad_plane_stress_pressure, soil_column_equaldof,
                       soil_structure_block, tie)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12
BEAM_RTOL = 1e-10


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_device_vs_golden_reference_vectors(name):
    mk, numberer, soe, _ = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = mk()
    D = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
    nd = ele_nd(spec)
    # the force-based beam iterates to |dW| < 1e-12 per element: its converged state, and with it
    # K and R, is reproducible to the iteration tolerance, not to the last bit
    tol = BEAM_RTOL if nd in (6, 12) else RTOL
    for s in range(NSTEPS):
        D.set_trial_disp(g[f"u{s}"]); D.update(); D.apply_load(0.25 * (s + 1))
        A, B = D.form_tangent(), D.form_unbalance()
        assert relerr(A, g[f"A{s}"]) < tol
        assert relerr(B, g[f"B{s}"]) < tol
        for e in range(len(g[f"K{s}"])):
            assert relerr(D.element_tangent(e, nd), g[f"K{s}"][e]) < tol
            assert relerr(D.element_resid(e, nd), g[f"R{s}"][e]) < tol
        D.commit()


@pytest.mark.parametrize("mat", [J2_STEEL, ELASTIC], ids=["j2", "elastic"])
@pytest.mark.parametrize("shape", ["brick", "quad"])
@pytest.mark.parametrize("numberer,soe", [(0, 0), (1, 1)])
def test_device_vs_oracle_load_history(mat, shape, numberer, soe):
    """several update/commit cycles with growing random displacements (well past yield)"""
    rng = np.random.default_rng(42)
    spec = (brick_block(5, 4, 3, mat=mat, distort=0.25, seed=3, body=(0.01, 0.0, -0.02)) if shape == "brick"
            else quad_plane(9, 6, mat=mat, lx=9.0, ly=6.0, distort=0.25, seed=4, body=(0.0, -0.03)))
    nd, order = (24, 6) if shape == "brick" else (8, 3)
    O = OracleBackend(spec, numberer, soe)
    D = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
    ids = O.ids()
    # untouched model: elastic tangent, zero residual
    assert relerr(D.form_tangent(), O.form_tangent()) < RTOL
    for s in range(5):
        u = rng.normal(0, 1.5e-3 * (s + 1), (spec.nn, spec.ndf)); u[ids < 0] = 0
        O.set_trial_disp(u); D.set_trial_disp(u); D.update()
        O.apply_load(0.2 * s); D.apply_load(0.2 * s)
        assert relerr(D.form_tangent(), O.form_tangent()) < RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_tangent(e, nd), O.ele_tangent(e, nd)) < RTOL
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < RTOL
        if s % 2 == 0:
            O.commit(); D.commit()


@pytest.mark.parametrize("shape", ["soilcolumn", "brick", "brick_elastic", "frame2d"])
@pytest.mark.parametrize("numberer,soe", [(0, 0), (1, 1), (1, 0)])
def test_device_vs_oracle_load_history_equaldof(shape, numberer, soe):
    """`equalDOF` (MP_Constraint, PlainHandler's -4 ids): constrained dofs share the retained dof's equation.  Rows fed
    by several nodes, two dofs of one element on one equation, ties onto fixed dofs; A and B against the oracle over a
    load history with commits, plus the nodal-mass terms of a transient tangent on the shared rows."""
    rng = np.random.default_rng(43)
    if shape == "soilcolumn":
        spec, sc = soil_column_equaldof(40, distort=0.2, seed=31), 2e-3
    elif shape == "brick":
        spec, sc = brick_periodic_equaldof(5, 4, 3, seed=32), 1.5e-3
    elif shape == "brick_elastic":
        spec, sc = brick_periodic_equaldof(3, 5, 4, mat=ELASTIC, dofs=(0, 1, 2), seed=33), 1.5e-3
    else:
        spec, sc = frame2d_diaphragm_equaldof(3, 4, 2), np.array((0.006, 0.003, 6e-5))
    beam = spec.groups[0].kind in (2, 3)
    tol = BEAM_RTOL if beam else RTOL
    mass = rng.uniform(0.01, 0.1, (spec.nn, spec.ndf))
    O = OracleBackend(spec, numberer, soe); O.set_mass(spec.node_tags, mass)
    D = xb.DeviceModel.from_spec(spec, setup=False); D.set_mass(spec.node_tags, mass); D.setup(numberer, soe); D.to_device(0)
    ids = O.ids()
    assert np.array_equal(D.ids(), ids)
    assert (np.bincount(ids[ids >= 0]) > 1).sum() >= 4          # equations really are shared
    assert relerr(D.form_tangent(), O.form_tangent()) < tol
    for s in range(5):
        u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * (s + 1); u[ids < 0] = 0
        tie(spec, u)
        O.set_trial_disp(u); D.set_trial_disp(u); D.update()
        O.apply_load(0.2 * s); D.apply_load(0.2 * s)
        if s == 3:      # Newmark terms: c1 K + c3 M (+ alphaM c2 M), P - M a - alphaM M v - R
            v, a = rng.normal(0, 1.0, (2, spec.nn, spec.ndf)); v[ids < 0] = 0; a[ids < 0] = 0
            tie(spec, v); tie(spec, a)
            for m in (O, D):
                m.set_rayleigh(0.3, 0.0, 0.0, 0.0); m.set_transient(1.0, 50.0, 1.0e4); m.set_vel_accel(v, a)
        assert relerr(D.form_tangent(), O.form_tangent()) < tol
        assert relerr(D.form_unbalance(), O.form_unbalance()) < tol
        if s % 2 == 0:
            O.commit(); D.commit()


def _newton(model, solve, nsteps, dlam, tol, max_iter, is_dev):
    """BasicAnalysisBuilder::analyzeStatic with LoadControl + NewtonRaphson + CTestNormDispIncr
    (newStep / solveCurrentStep / commit), the linear solve delegated to `solve`."""
    hist = []
    lam = 0.0
    for _ in range(nsteps):
        lam += dlam
        model.apply_load(lam)                      # LoadControl::newStep: no state determination here
        B = model.form_unbalance()
        norms = []
        for it in range(max_iter):
            A = model.form_tangent()
            dU = solve(A, B)
            if is_dev:
                model.incr_trial_disp(dU); model.update()
            else:
                u = model._u; ids = model.ids()
                u[ids >= 0] += dU[ids[ids >= 0]]
                model.set_trial_disp(u)
            B = model.form_unbalance()
            norms.append(float(np.linalg.norm(dU)))
            if norms[-1] <= tol:
                break
        hist.append(norms)
        model.commit()
    return hist


def _newton_counts_match(spec, nsteps, min_iters):
    """Static Newton (LoadControl, NormDispIncr) driven once by the oracle and once by the device:
    identical iteration counts per step and the same convergence history.  The tolerance is picked
    from a fixed list so that, in the oracle's own history, no deciding norm sits within 3x of it:
    an iteration count must not hinge on the last bits of a norm."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    O = OracleBackend(spec, 1, 1)
    ptr, idx = O.csr()
    neq = O.neq

    def solve(A, B):
        return spla.spsolve(sp.csr_matrix((A, idx, ptr), shape=(neq, neq)).tocsc(), B)

    best = None
    for tol in (1e-6, 3e-7, 1e-7, 3e-8, 1e-8, 3e-9, 1e-9):
        O = OracleBackend(spec, 1, 1); O._u = np.zeros((spec.nn, spec.ndf))
        h = _newton(O, solve, nsteps, 1.0, tol, 25, False)
        margin = min(min(x[-2] / tol, tol / max(x[-1], 1e-300)) for x in h)
        if best is None or margin > best[0]:
            best = (margin, tol, h, O._u.copy())
    margin, tol, ho, uo = best
    assert margin >= 3.0, (margin, tol)
    D = xb.DeviceModel.from_spec(spec, 1, 1).to_device(0)
    hd = _newton(D, solve, nsteps, 1.0, tol, 25, True)
    assert [len(h) for h in ho] == [len(h) for h in hd]          # identical iteration counts
    assert max(len(h) for h in ho) >= min_iters                    # the steps really go inelastic
    for a, b in zip(ho, hd):
        assert np.allclose(a[:-1], b[:-1], rtol=1e-5, atol=1e-11)  # same convergence history
    assert relerr(D.trial_disp(), uo) < 1e-8


@pytest.mark.parametrize("shape", ["brick", "quad", "frame", "frame3d", "soilcolumn_equaldof", "frame_equaldof"])
def test_newton_iteration_counts_match_oracle(shape):
    if shape == "soilcolumn_equaldof":      # sheared soil column with tied (periodic) sides
        spec = soil_column_equaldof(12, mat=J2_STEEL, distort=0.1)
        spec.loads = np.array([[1 + 2 * 12, 22.0, -3.0], [1 + 2 * 6, 10.0, 0.0]])
        _newton_counts_match(spec, 8, 6)
    elif shape == "frame_equaldof":         # RC frame with rigid-diaphragm ties
        _newton_counts_match(frame2d_diaphragm_equaldof(2, 3, 2, lateral=22.0, gravity=-40.0), 5, 4)
    elif shape == "brick":
        spec = brick_block(4, 4, 6, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.2, 0.0, -0.5))
        _newton_counts_match(spec, 8, 6)
    elif shape == "quad":
        spec = quad_plane(16, 4, mat=J2_STEEL, lx=8.0, ly=2.0)
        spec.loads[:, 1:] = [0.0, -10.0]
        _newton_counts_match(spec, 8, 6)
    elif shape == "frame":   # load-controlled pushover of the RC frame (forceBeamColumn, fibre sections)
        _newton_counts_match(frame2d(2, 3, 2, lateral=22.0, gravity=-40.0), 5, 5)
    else:   # 3D space frame (ForceBeamColumn3d, FiberSection3d): biaxial push at a roof corner
        _newton_counts_match(frame3d(1, 1, 2, ndiv=2, lateral=(20.0, 12.0), gravity=-40.0), 5, 5)


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape,numberer,soe", [("brick", 1, 0), ("quad", 0, 1), ("mixed", 1, 1), ("soilcolumn_equaldof", 1, 0), ("frame2d_gravity", 1, 0), ("soil_frame_mixed_ndf", 1, 0),
                                                ("soilcolumn_equaldof", 0, 1), ("frame2d_pdelta", 1, 0), ("frame3d_pdelta", 1, 0), ("frame3d_eleloads", 1, 0), ("frame2d_legendre", 1, 0), ("frame3d_radau", 1, 0), ("frame2d_concrete01", 1, 0), ("frame2d_jntoffset", 1, 0), ("frame3d_jntoffset", 1, 0), ("frame2d_corot", 1, 0), ("frame2d_partial_load", 1, 0), ("frame3d_partial_load", 1, 0), ("frame2d_elasticpp", 1, 0),
                                                ("soilcolumn_equaldof@T", 1, 0), ("brick@T", 0, 1), ("quad@T", 1, 0), ("mixed@T", 1, 1), ("frame2d_gravity@T", 1, 0), ("frame3d_pdelta@T", 1, 0), ("soil_frame_mixed_ndf@T", 1, 0)])
def test_reference_newton_loop_drives_device_path(shape, numberer, soe):
    """The drop-in, end to end: the REFERENCE'S OWN StaticAnalysis objects (AnalysisModel, PlainHandler, numberer,
    SparseGenCol/Row SOE and solver, NewtonRaphson, CTestNormDispIncr, LoadControl::newStep) run a load-controlled
    Newton analysis twice on the same Domain description -- once unmodified on the CPU, once with the integrator
    of INTEGRATION.md (oracle/ref_glue.cpp) that reads the model out of the Domain and routes formTangent /
    formUnbalance / update / commit through the C ABI to the device.  Same iteration count on every step (decided
    by the reference's own convergence test), same norms, same final displacements."""
    from modelspec import GLUE_SO, RefBackend
    # "...@T": the same under `constraints Transformation` on both sides -- the handler numbers and assembles these models as
    # PlainHandler does but updates the elements next to constrained nodes once more at every applyLoad, which changes the
    # Newton history once the model yields (tests/test_oracle.py); the binding switches the device's second update on
    handler = 1 if shape.endswith("@T") else 0
    shape = shape.split("@")[0]
    if shape == "brick":
        mk = lambda: brick_block(4, 4, 6, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.2, 0.0, -0.5))
    elif shape == "quad":
        def mk():
            sp = quad_plane(16, 4, mat=J2_STEEL, lx=8.0, ly=2.0); sp.loads[:, 1:] = [0.0, -10.0]; return sp
    elif shape == "soil_frame_mixed_ndf":
        def mk():   # quads on 2-dof nodes + force beams on 3-dof nodes, read out of the reference's Domain node by node
            from modelspec import soil_frame_2d
            sp = soil_frame_2d(lateral=35.0, gravity=-80.0); return sp
    elif shape == "frame2d_gravity":
        def mk():   # `eleLoad -beamUniform` on the girders read out of the load pattern, pushed well into the inelastic range
            from modelspec import with_beam_gravity
            sp = with_beam_gravity(frame2d(2, 3, 2, lateral=15.0), w=-0.2, seed=1); return sp
    elif shape in ("frame2d_partial_load", "frame3d_partial_load"):
        def mk():   # trapezoidal `eleLoad -beamUniform ... aOverL bOverL` (Beam2d / Beam3dPartialUniformLoad) read out of the load pattern
            from modelspec import with_beam_partial_loads
            return (with_beam_partial_loads(frame2d(2, 3, 2, lateral=15.0), w=-0.3, seed=4) if shape == "frame2d_partial_load"
                    else with_beam_partial_loads(frame3d(1, 1, 2, ndiv=2, lateral=(14.0, 8.0)), w=-0.1, seed=4))
    elif shape == "frame3d_eleloads":
        def mk():   # `eleLoad -beamUniform` and `-beamPoint` (Beam3dUniformLoad, Beam3dPointLoad) read out of the load pattern
            from modelspec import with_beam_gravity, with_beam_point_loads
            return with_beam_point_loads(with_beam_gravity(frame3d(1, 1, 2, ndiv=2, lateral=(14.0, 8.0)), w=-0.06, seed=1), P=-2.5, seed=2)
    elif shape in ("frame2d_jntoffset", "frame3d_jntoffset"):
        def mk():   # rigid end zones under geomTransf PDelta: nodeIOffset / nodeJOffset read out of the elements' CrdTransf
            from modelspec import with_joint_offsets, with_pdelta
            return with_pdelta(with_joint_offsets(frame2d(2, 3, 2, lateral=15.0, gravity=-150.0) if shape == "frame2d_jntoffset"
                                                  else frame3d(1, 1, 2, ndiv=2, lateral=(18.0, 10.0), gravity=-90.0), seed=3))
    elif shape == "frame2d_corot":
        def mk():   # `geomTransf Corotational` (CorotCrdTransf2d) told from the others by dynamic_cast: heavy gravity, then the push
            from modelspec import with_corot
            return with_corot(frame2d(2, 3, 2, lateral=15.0, gravity=-150.0))
    elif shape == "frame2d_elasticpp":
        def mk():   # ElasticPP bars read out of the reference's FiberSection2d (E, fyp / E, fyn / E, ezero)
            from modelspec import steel01_elastic_frame
            return steel01_elastic_frame(2, "elasticpp")
    elif shape == "frame2d_concrete01":
        def mk():   # Concrete01 core, Steel01 bars, bilinear Elastic cover read out of the reference's FiberSection2d
            from modelspec import steel01_elastic_frame
            return steel01_elastic_frame(2)
    elif shape in ("frame2d_legendre", "frame3d_radau"):
        def mk():   # -integration Legendre / Radau: the binding reads the section locations and weights out of the element's BeamIntegration
            from modelspec import with_beam_integration
            return (with_beam_integration(frame2d(2, 3, 2, nip=4, lateral=28.0, gravity=-60.0), 1) if shape == "frame2d_legendre"
                    else with_beam_integration(frame3d(1, 1, 2, ndiv=2, nip=5, lateral=(25.0, 15.0), gravity=-40.0), 2))
    elif shape in ("frame2d_pdelta", "frame3d_pdelta"):
        def mk():   # `geomTransf PDelta` read out of the elements' CrdTransf: heavy gravity, then the lateral push
            from modelspec import with_pdelta
            return with_pdelta(frame2d(2, 3, 2, lateral=15.0, gravity=-150.0) if shape == "frame2d_pdelta"
                               else frame3d(1, 1, 2, ndiv=2, lateral=(18.0, 10.0), gravity=-90.0))
    elif shape == "soilcolumn_equaldof":
        def mk():   # MP_Constraints read out of the Domain (`equalDOF`): sheared soil column with tied sides
            sp = soil_column_equaldof(12, mat=J2_STEEL, distort=0.1)
            sp.loads = np.array([[1 + 2 * 12, 22.0 * 8, -3.0 * 8], [1 + 2 * 6, 10.0 * 8, 0.0]]); return sp
    else:
        def mk():   # soil (J2) + footing (elastic): two element batches, loaded well past first yield
            sp = soil_structure_block(5, 5, 5, distort=0.1, seed=2); sp.loads[:, 1:] *= 12.0; return sp
    nsteps, dl, max_iter = 8, 1.0 / 8, 25
    best = None
    for tol in (1e-6, 1e-7, 1e-8, 1e-9):          # a tolerance no deciding norm sits within 3x of
        C = RefBackend(mk(), numberer, soe, dlambda=dl, test=0, tol=tol, max_iter=max_iter, handler=handler)
        rc, iters, norms = C.analyze_static(nsteps)
        assert rc == 0
        margin = min(min(norms[s, iters[s] - 2] / tol if iters[s] > 1 else 1e9, tol / max(norms[s, iters[s] - 1], 1e-300)) for s in range(nsteps))
        if best is None or margin > best[0]:
            best = (margin, tol, iters.copy(), norms.copy(), C.get_trial_disp())
    margin, tol, it_cpu, nm_cpu, u_cpu = best
    assert margin >= 3.0 and it_cpu.max() >= 4
    D = RefBackend(mk(), defer_setup=True, so=GLUE_SO, handler=handler)
    D.setup_glue_loadcontrol(numberer, soe, dl, test=0, tol=tol, max_iter=max_iter)
    rc, it_dev, nm_dev = D.analyze_static(nsteps)
    assert rc == 0
    assert it_dev.tolist() == it_cpu.tolist()                       # identical Newton iteration counts
    for s in range(nsteps):
        assert np.allclose(nm_dev[s, :it_dev[s] - 1], nm_cpu[s, :it_cpu[s] - 1], rtol=1e-5, atol=1e-11)
    assert relerr(D.glue_trial_disp(), u_cpu) < 1e-8
    calls, launches = D.glue_counts()
    assert calls[0] == it_cpu.sum() and calls[3] == nsteps and launches > 0     # the device did the work


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_failed_step_reverts_the_device_state():
    """A step that does not converge (3 iterations allowed, the plastic steps need more): the reference's analysis loop
    calls Domain::revertToLastCommit and the integrator's revertToLastStep -- same failing step, same iteration counts,
    and the DEVICE state is back at the last commit, like the reference's Domain"""
    from modelspec import GLUE_SO, RefBackend
    mk = lambda: brick_block(4, 4, 6, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.2, 0.0, -0.5))
    C = RefBackend(mk(), 1, 0, dlambda=1.0 / 8, test=0, tol=1e-8, max_iter=3)
    rc_cpu, it_cpu, _ = C.analyze_static(8)
    assert rc_cpu == -3 and it_cpu[:5].tolist() == [2, 2, 2, 2, 2]          # steps 1-5 converge, step 6 fails
    D = RefBackend(mk(), defer_setup=True, so=GLUE_SO)
    D.setup_glue_loadcontrol(1, 0, 1.0 / 8, test=0, tol=1e-8, max_iter=3)
    rc_dev, it_dev, _ = D.analyze_static(8)
    assert rc_dev == -3 and it_dev.tolist() == it_cpu.tolist()
    u_cpu = C.get_trial_disp()                                               # the state of the last commit (step 5)
    assert np.abs(u_cpu).max() > 1e-3 and relerr(D.glue_trial_disp(), u_cpu) < 1e-9


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_loop_reads_plane_stress_and_pressure_out_of_the_domain():
    """FourNodeQuad with the PlaneStress material copy and a surface pressure: the glue reads both out of the reference's
    Domain (theMaterial[0]->getType(), pressure); the reference's own analysis on CPU and on the device path agree."""
    from modelspec import GLUE_SO, RefBackend
    mk = lambda: quad_plane_stress_pressure(10, 6, 1, 4.0)
    C = RefBackend(mk(), 1, 0, dlambda=0.5, test=0, tol=1e-9, max_iter=10)
    rc, it_cpu, nm_cpu = C.analyze_static(2)
    assert rc == 0
    D = RefBackend(mk(), defer_setup=True, so=GLUE_SO)
    D.setup_glue_loadcontrol(1, 0, 0.5, test=0, tol=1e-9, max_iter=10)
    rc, it_dev, nm_dev = D.analyze_static(2)
    assert rc == 0 and it_dev.tolist() == it_cpu.tolist()
    u_cpu = C.get_trial_disp()
    assert np.abs(u_cpu).max() > 1e-3 and relerr(D.glue_trial_disp(), u_cpu) < 1e-9


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", ["brick", "quad", "brick@T", "quad@T"])
def test_reference_newmark_loop_drives_device_path(shape):
    """The transient drop-in: the reference's own Newmark bookkeeping (U, Udot, Udotdot, predictor), NewtonRaphson and
    convergence test, with newStep / update / formTangent / formUnbalance / commit of the integrator routed to the
    device (oracle/ref_glue.cpp, B200Newmark).  Nodal masses, `rayleigh` factors and the material density are read
    out of the reference's Domain.  Same iteration counts, norms, displacements, velocities and accelerations as
    the unmodified reference run."""
    from golden_cases import J2_STEEL_RHO, RAYLEIGH
    from modelspec import GLUE_SO, RefBackend
    handler = 1 if shape.endswith("@T") else 0          # "...@T": `constraints Transformation` on both sides
    shape = shape.split("@")[0]
    if shape == "brick":
        mk = lambda: brick_block(3, 3, 4, mat=J2_STEEL_RHO, lz=3.0, load=(240.0, 0.0, -30.0), distort=0.15, seed=3)
    else:
        def mk():
            sp = quad_plane(12, 4, mat=J2_STEEL_RHO, lx=6.0, ly=2.0, distort=0.1, seed=4); sp.loads[:, 1:] = [0.0, -260.0]; return sp
    nsteps, dt, gamma, beta, max_iter = 8, 0.02, 0.5, 0.25, 25

    def build(so=None):
        spec = mk()
        R = RefBackend(spec, defer_setup=True, so=so, handler=handler)
        mass = np.zeros((spec.nn, spec.ndf)); mass[:] = 0.05
        R.set_mass(spec.node_tags, mass); R.set_rayleigh(*RAYLEIGH)
        return R

    # stiffness-proportional damping on the CURRENT tangent makes Newton converge linearly here (6-8 iterations a
    # step): pick the tolerance that no deciding norm sits close to (the two runs agree to ~1e-6 in every norm)
    best = None
    for tol in (1e-8, 3e-9, 1e-9, 3e-10, 1e-10):
        C = build(); C.setup_transient(1, 1, gamma, beta, test=0, tol=tol, max_iter=max_iter)
        rc, it_cpu, nm_cpu = C.analyze_transient(nsteps, dt)
        assert rc == 0
        margin = min(min(nm_cpu[s, it_cpu[s] - 2] / tol if it_cpu[s] > 1 else 1e9, tol / max(nm_cpu[s, it_cpu[s] - 1], 1e-300)) for s in range(nsteps))
        if best is None or margin > best[0]:
            best = (margin, tol, it_cpu.copy(), nm_cpu.copy(), C)
    margin, tol, it_cpu, nm_cpu, C = best
    assert margin >= 1.3 and it_cpu.max() >= 3, (margin, tol)
    D = build(GLUE_SO); D.setup_glue_newmark(1, 1, gamma, beta, test=0, tol=tol, max_iter=max_iter)
    rc, it_dev, nm_dev = D.analyze_transient(nsteps, dt)
    assert rc == 0
    assert it_dev.tolist() == it_cpu.tolist()
    for s in range(nsteps):
        assert np.allclose(nm_dev[s, :it_dev[s] - 1], nm_cpu[s, :it_cpu[s] - 1], rtol=1e-5, atol=1e-12)
    assert relerr(D.glue_trial_disp(), C.get_trial_disp()) < 1e-8
    vc, ac = C.vel_accel(); vd, ad = D.vel_accel()       # the reference's nodes, driven by the device path's solution
    assert relerr(vd, vc) < 1e-7 and relerr(ad, ac) < 1e-7
    calls, launches = D.glue_counts()
    assert calls[0] == it_cpu.sum() and calls[3] == nsteps and launches > 0


@pytest.mark.skipif(not have_glue(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["dc_cantilever_fiber", "dc_frame3d", "dc_brick_j2"])
def test_reference_displacement_control_drives_device_path(name):
    """BASELINE configs[0] (the Ex2b cantilever pushover with the RC fibre section), the 3D space-frame pushover and
    the J2 brick column: the reference's own DisplacementControl + NewtonRaphson + CTestNormDispIncr objects, with the
    integrator's model-facing calls routed to the device (B200DisplacementControl in oracle/ref_glue.cpp; frames, fibre
    sections and uniaxial materials read out of the reference's Domain).  Checked against the history the UNMODIFIED
    reference produced (tests/golden/dc_*.npz): identical iteration counts, load factors, displacements."""
    from golden_cases import DISPCONTROL_CASES
    from modelspec import GLUE_SO, RefBackend
    mk, numberer, soe, node, dof, incr, nsteps, tol, max_iter = DISPCONTROL_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    D = RefBackend(mk(), defer_setup=True, so=GLUE_SO)
    D.setup_glue_dispcontrol(numberer, soe, int(g["node"]), dof, incr, test=0, tol=tol, max_iter=max_iter)
    rc, iters, norms, lam = D.analyze_static_lam(nsteps)
    assert rc == 0
    assert iters.tolist() == g["iters"].tolist()
    assert relerr(lam, g["lam"]) < 1e-8
    assert relerr(D.glue_trial_disp(), g["u"]) < 1e-8
    for s in range(nsteps):
        assert np.allclose(norms[s, :iters[s] - 1], g["norms"][s, :iters[s] - 1], rtol=1e-5, atol=1e-12)
    calls, launches = D.glue_counts()
    assert calls[3] == nsteps and launches > 0


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["dc_cantilever_fiber", "dc_frame3d", "dc_brick_j2"])
def test_reference_displacement_control_under_transformation_handler(name):
    """`constraints Transformation` + `integrator DisplacementControl`: here the reference calls applyLoadDomain -- and with it
    the handler's second update of the elements with a constrained node -- in EVERY iteration, between incrDisp and
    updateDomain (DisplacementControl.cpp:121,210): a force beam at the fixed base is updated twice per iteration with the
    same increment.  The unmodified reference on the CPU against the same loop over the device path
    (`constraints_transformation`): same iteration counts, load factors, displacements."""
    from golden_cases import DISPCONTROL_CASES, control_node
    from modelspec import GLUE_SO, RefBackend
    mk, numberer, soe, node, dof, incr, nsteps, tol, max_iter = DISPCONTROL_CASES[name]
    nd = control_node(mk(), node)
    C = RefBackend(mk(), defer_setup=True, handler=1)
    C.setup_dispcontrol(numberer, soe, nd, dof, incr, test=0, tol=tol, max_iter=max_iter)
    rc, it_cpu, nm_cpu, lam_cpu = C.analyze_static_lam(nsteps)
    assert rc == 0
    D = RefBackend(mk(), defer_setup=True, so=GLUE_SO, handler=1)
    D.setup_glue_dispcontrol(numberer, soe, nd, dof, incr, test=0, tol=tol, max_iter=max_iter)
    rc, it_dev, nm_dev, lam_dev = D.analyze_static_lam(nsteps)
    assert rc == 0
    assert it_dev.tolist() == it_cpu.tolist()
    assert relerr(lam_dev, lam_cpu) < 1e-8 and relerr(D.glue_trial_disp(), C.get_trial_disp()) < 1e-8
    for s in range(nsteps):
        assert np.allclose(nm_dev[s, :it_dev[s] - 1], nm_cpu[s, :it_cpu[s] - 1], rtol=1e-5, atol=1e-12)


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", ["frame2d", "frame3d", "frame2d_pdelta", "frame3d_pdelta", "frame2d_rho", "frame2d@T", "frame3d_pdelta@T"])
def test_reference_newmark_loop_drives_device_frames(shape):
    """BASELINE configs[3] in small: RC frames of forceBeamColumn elements (Steel02 / Concrete02 fibre sections),
    transient Newmark with nodal masses and Rayleigh damping, run by the reference's own objects on the CPU and
    with the device-backed integrator: same iteration counts and responses."""
    from golden_cases import RAYLEIGH
    from modelspec import GLUE_SO, RefBackend
    from modelspec import with_pdelta
    handler = 1 if shape.endswith("@T") else 0          # "...@T": `constraints Transformation` on both sides
    shape = shape.split("@")[0]
    if shape == "frame2d": mk = lambda: frame2d(2, 2, 2, lateral=30.0)
    elif shape == "frame3d": mk = lambda: frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0))
    elif shape == "frame2d_rho":               # forceBeamColumn -mass: the glue reads rho out of the elements
        from modelspec import with_beam_rho
        mk = lambda: with_beam_rho(frame2d(2, 2, 2, lateral=30.0), 2.0e-3)
    elif shape == "frame2d_pdelta": mk = lambda: with_pdelta(frame2d(2, 2, 2, lateral=30.0, gravity=-150.0))      # geomTransf PDelta + rayleigh
    else: mk = lambda: with_pdelta(frame3d(1, 1, 2, ndiv=2, lateral=(25.0, 15.0), gravity=-90.0))
    nsteps, dt, gamma, beta, max_iter = 6, 0.02, 0.5, 0.25, 25

    def build(so=None):
        spec = mk()
        R = RefBackend(spec, defer_setup=True, so=so, handler=handler)
        mass = np.zeros((spec.nn, spec.ndf)); mass[:, :spec.ndm] = 0.05
        R.set_mass(spec.node_tags, mass); R.set_rayleigh(*RAYLEIGH)
        return R

    best = None
    for tol in (1e-7, 1e-8, 1e-9, 1e-10):
        C = build(); C.setup_transient(1, 1, gamma, beta, test=0, tol=tol, max_iter=max_iter)
        rc, it_cpu, nm_cpu = C.analyze_transient(nsteps, dt)
        assert rc == 0
        margin = min(min(nm_cpu[s, it_cpu[s] - 2] / tol if it_cpu[s] > 1 else 1e9, tol / max(nm_cpu[s, it_cpu[s] - 1], 1e-300)) for s in range(nsteps))
        if best is None or margin > best[0]:
            best = (margin, tol, it_cpu.copy(), nm_cpu.copy(), C)
    margin, tol, it_cpu, nm_cpu, C = best
    assert margin >= 1.5, (margin, tol)
    D = build(GLUE_SO); D.setup_glue_newmark(1, 1, gamma, beta, test=0, tol=tol, max_iter=max_iter)
    rc, it_dev, nm_dev = D.analyze_transient(nsteps, dt)
    assert rc == 0
    assert it_dev.tolist() == it_cpu.tolist()
    assert relerr(D.glue_trial_disp(), C.get_trial_disp()) < 1e-7
    vc, ac = C.vel_accel(); vd, ad = D.vel_accel()
    assert relerr(vd, vc) < 1e-6 and relerr(ad, ac) < 1e-6


def test_revert_to_last_commit_and_incr():
    rng = np.random.default_rng(0)
    spec = brick_block(3, 3, 3, distort=0.1)
    O = OracleBackend(spec, 0, 1); D = xb.DeviceModel.from_spec(spec, 0, 1).to_device(0)
    ids = O.ids()
    u1 = rng.normal(0, 4e-3, (spec.nn, 3)); u1[ids < 0] = 0
    O.set_trial_disp(u1); D.set_trial_disp(u1); D.update(); O.apply_load(0.5); D.apply_load(0.5); O.commit(); D.commit()
    u2 = u1 + rng.normal(0, 4e-3, (spec.nn, 3)); u2[ids < 0] = 0
    O.set_trial_disp(u2); D.set_trial_disp(u2); D.update(); O.apply_load(0.9); D.apply_load(0.9)
    O.revert(); D.revert_to_last_commit()            # Domain::revertToLastCommit: also the committed load factor (0.5)
    assert relerr(D.trial_disp(), u1) == 0.0
    assert relerr(D.form_unbalance(), O.form_unbalance()) < RTOL
    dU = rng.normal(0, 1e-3, O.neq)
    D.incr_trial_disp(dU)
    u3 = u1.copy(); u3[ids >= 0] += dU[ids[ids >= 0]]
    assert relerr(D.trial_disp(), u3) < 1e-15


@pytest.mark.parametrize("seed", range(6))
def test_ragged_random_meshes_device_vs_oracle(seed):
    """the ragged meshes of tests/test_host_setup.py (holes, nodes without elements, disconnected pieces, extra fixes,
    `equalDOF` ties, shuffled elements) through the device path: A and B against the oracle over a short history"""
    from test_host_setup import ragged_spec
    rng = np.random.default_rng(500 + seed)
    spec = ragged_spec(seed)
    for numberer, soe in ((0, 0), (1, 1)):
        O = OracleBackend(spec, numberer, soe); D = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
        ids = O.ids()
        for s in range(3):
            u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 3)); u[ids < 0] = 0
            tie(spec, u)
            O.set_trial_disp(u); D.set_trial_disp(u); D.update()
            O.apply_load(0.3 * s); D.apply_load(0.3 * s)
            assert relerr(D.form_tangent(), O.form_tangent()) < RTOL
            assert relerr(D.form_unbalance(), O.form_unbalance()) < RTOL
            O.commit(); D.commit()


@pytest.mark.parametrize("shape", ["brick", "quad_ps_j2", "frame2d", "frame3d"])
def test_revert_to_start(shape):
    """Domain::revertToStart (the `reset` command): after a committed plastic history the model is back at its initial
    state (bitwise the initial A), and the next step matches the oracle (which is pinned to the reference's revertToStart)"""
    rng = np.random.default_rng(2)
    spec, sc = {"brick": (brick_block(3, 3, 3, distort=0.1), 4e-3),
                "quad_ps_j2": (quad_plane_stress_pressure(5, 4, 1, 1.5, mat=J2_STEEL), 3e-3),
                "frame2d": (frame2d(2, 2, 2), np.array((0.006, 0.003, 6e-5))),
                "frame3d": (frame3d(1, 1, 2), np.array((0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4)))}[shape]
    tol = BEAM_RTOL if shape.startswith("frame") else RTOL
    O = OracleBackend(spec, 1, 1); D = xb.DeviceModel.from_spec(spec, 1, 1).to_device(0)
    ids = O.ids()
    A0, B0 = D.form_tangent().copy(), D.form_unbalance().copy()
    for s in range(3):
        u = rng.normal(0, 1, (spec.nn, spec.ndf)) * sc * (s + 1); u[ids < 0] = 0
        O.set_trial_disp(u); D.set_trial_disp(u); D.update()
        O.apply_load(0.4 * (s + 1)); D.apply_load(0.4 * (s + 1)); O.commit(); D.commit()
    assert relerr(D.form_tangent(), A0) > 1e-3                     # the history left its mark
    O.revert_to_start(); D.revert_to_start()
    assert np.array_equal(D.form_tangent(), A0) and np.array_equal(D.form_unbalance(), B0)
    assert relerr(D.form_tangent(), O.form_tangent()) < tol and np.abs(D.trial_disp()).max() == 0.0
    u = rng.normal(0, 1, (spec.nn, spec.ndf)) * sc; u[ids < 0] = 0
    O.set_trial_disp(u); D.set_trial_disp(u); D.update(); O.apply_load(0.3); D.apply_load(0.3)
    assert relerr(D.form_tangent(), O.form_tangent()) < tol
    assert relerr(D.form_unbalance(), O.form_unbalance()) < tol


def test_j2_plane_stress_quads_history_and_revert():
    """FourNodeQuad with J2Plasticity's PlaneStress copy (J2PlaneStress): the out-of-plane strain is a state of its own --
    every trial starts from the LAST TRIAL's value, commit stores it, revertToLastCommit restores it -- and the
    iteration on sigma_22 = 0 stops at 1e-8 sigma_0, so device and oracle agree to 1e-12 only if they make the same
    sequence of integrator calls.  Uncommitted trials, commits and a revert in between."""
    rng = np.random.default_rng(9)
    spec = quad_plane_stress_pressure(7, 5, 1, 1.5, mat=J2_STEEL, seed=37)
    O = OracleBackend(spec, 1, 0); D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    assert relerr(D.form_tangent(), O.form_tangent()) < RTOL          # condensed elastic tangent of the untouched model
    for s in range(6):
        u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 2)); u[ids < 0] = 0
        O.set_trial_disp(u); D.set_trial_disp(u); D.update()
        O.apply_load(0.2 * s); D.apply_load(0.2 * s)
        assert relerr(D.form_tangent(), O.form_tangent()) < RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < RTOL
        for e in (0, O.ne - 1):
            assert relerr(D.element_tangent(e, 8), O.ele_tangent(e, 8)) < RTOL
        if s in (1, 4):
            O.commit(); D.commit()
        if s == 3:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < RTOL
            assert relerr(D.form_unbalance(), O.form_unbalance()) < RTOL
    sg, tg = D.gp_response(3, 2, 3)                                   # the condensed tangent: symmetric, plane-stress soft
    assert np.allclose(tg, tg.T, rtol=0, atol=1e-9 * np.abs(tg).max()) and np.isfinite(sg).all()


def test_full_size_properties_brick():
    """A 200k-element J2 block (too large for the oracle in seconds): properties that do not
    depend on size -- symmetry of A, rigid-body null space of the elastic operator, row sums of
    B equal to the applied load, determinism (bitwise identical A on a second pass)."""
    import scipy.sparse as sp
    n = 58
    spec = brick_block(n, n, n, mat=J2_STEEL, fix_face="z0")
    D = xb.DeviceModel.from_spec(spec, 0, 1).to_device(0)
    ids = D.ids()
    ptr, idx = D.pattern()
    A0 = D.form_tangent()
    M = sp.csr_matrix((A0, idx, ptr), shape=(D.neq, D.neq))
    asym = abs(M - M.T).max() / abs(M).max()
    assert asym < 1e-13
    # elastic state: K * (uniform translation) must vanish on rows not coupled to the fixed base
    t = np.zeros((spec.nn, 3)); t[:, 0] = 1.0
    x = np.zeros(D.neq); x[ids[ids >= 0]] = t[ids >= 0]
    r = M @ x
    far = ids[spec.crd[:, 2] > 1.5 / n]
    far = far[far >= 0]
    assert np.abs(r[far]).max() < 1e-9 * abs(M).max()
    # plastic state + determinism
    rng = np.random.default_rng(1)
    u = rng.normal(0, 3e-3, (spec.nn, 3)); u[ids < 0] = 0
    D.set_trial_disp(u); D.update(); D.apply_load(1.0)
    A1 = D.form_tangent(); A2 = D.form_tangent()
    assert np.array_equal(A1, A2)
    # xb_form_tangent with a host destination forms A range by range and streams finished rows
    # out; the split calls assemble in one launch and copy once: same bits
    D.form_element_tangents(); A3 = np.empty(D.nnz); D.assemble_tangent(A3); D.synchronize()
    assert np.array_equal(A1, A3)
    assert np.abs(A1 - A0).max() > 1e-3 * np.abs(A0).max()
    M1 = sp.csr_matrix((A1, idx, ptr), shape=(D.neq, D.neq))
    assert abs(M1 - M1.T).max() / abs(M1).max() < 1e-13
    # B = lambda*P - sum R: with zero displacement the residual is the load itself
    D.set_trial_disp(np.zeros_like(u)); D.update()
    B = D.form_unbalance()
    P = np.zeros(D.neq); ld = spec.loads
    for row in ld:
        nidx = int(row[0]) - 1
        for d in range(3):
            if ids[nidx, d] >= 0:
                P[ids[nidx, d]] += row[1 + d]
    # committed plastic strains are still zero (nothing was committed), so R(0) = 0
    assert np.abs(B - P).max() < 1e-12 * max(np.abs(P).max(), 1.0)


@pytest.mark.parametrize("storage", [0, 1], ids=["rows", "records"])
@pytest.mark.parametrize("name", list(R1_CASES))
def test_tangent_and_unbalance_bits_equal_round1(name, storage):
    """formTangent of round 2 -- the tangent kernel's TMA output as node-major rows (streamed assembly) or as symmetric
    element records (gathered assembly) -- reproduces the round-1 device path BIT FOR BIT: same block values, same
    FE_Element order of additions.  The digests were taken on a B200 with the round-1 library (tests/golden/make_r1_bits.py)."""
    import json
    with open(os.path.join(GOLD, "r1_tangent_bits.json")) as f:
        want = json.load(f)[name]
    assert r1_run_case(xb, name, {"brick_storage": storage}) == want


def test_shuffled_tags_and_tangent_options_bitwise():
    """Element tags arrive shuffled; DOF numbers, pattern and FE order are the reference's all the same and A matches the
    oracle.  The options (ranged formTangent, hand-tuned or generic gathered assembly, rows or records) do not change a bit."""
    spec = brick_block(48, 40, 36, mat=J2_STEEL, distort=0.2, seed=3)
    g = spec.groups[0]
    p = np.random.default_rng(0).permutation(len(g.tags))
    g.tags, g.conn, g.mat, g.par = g.tags[p], g.conn[p], g.mat[p], g.par[p]
    rng = np.random.default_rng(5)
    O = OracleBackend(spec, 1, 0)
    ids = O.ids()
    u = rng.normal(0, 2.5e-3, (spec.nn, 3)); u[ids < 0] = 0
    O.set_trial_disp(u); O.apply_load(0.7)
    Ao, Bo = O.form_tangent(), O.form_unbalance()
    res = {}
    for opt in ((1, 1, 1), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0), (0, 1, 0)):
        D = xb.DeviceModel.from_spec(spec, 1, 0, options={"brick_storage": opt[2]}).to_device(0)
        D.set_option("ranged_tangent", opt[0]).set_option("fast_assembly", opt[1])
        assert np.array_equal(D.ids(), ids) and np.array_equal(D.element_tags(), O.fe_ids(24)[0])
        D.set_trial_disp(u); D.update(); D.apply_load(0.7)
        A = D.form_tangent(); B = D.form_unbalance()
        assert relerr(A, Ao) < RTOL and relerr(B, Bo) < RTOL
        for e in (0, 1234, O.ne - 1):          # FE-order accessors see through the storage order
            assert relerr(D.element_tangent(e, 24), O.ele_tangent(e, 24)) < RTOL
            assert relerr(D.element_resid(e, 24), O.ele_resid(e, 24)) < RTOL
        D.form_element_tangents(); A3 = np.empty(D.nnz); D.assemble_tangent(A3); D.synchronize()
        assert np.array_equal(A, A3)
        A4 = D.form_tangent(host=False); D.synchronize()
        D.commit()
        D.set_trial_disp(1.5 * u); D.update()
        res[opt] = (A, B, D.form_tangent(host=True), D.form_unbalance())
    for opt in ((0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0), (0, 1, 0)):
        for x, y in zip(res[(1, 1, 1)], res[opt]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("soe", [2, 3, 4], ids=["BandGeneral", "ProfileSPD", "Umfpack"])
def test_band_profile_umfpack_storage_device_vs_oracle(soe):
    """`system BandGeneral | ProfileSPD | Umfpack`: the assembly kernels write A into the SOE's own array (LAPACK band,
    upper profile by columns, Umfpack's Ax) -- against the oracle's restatement of the SOEs' addA, which is pinned to the
    live BandGenLinSOE / ProfileSPDLinSOE (tests/test_oracle.py); bricks (record assembly), equalDOF (shared rows), quads
    and force beams, over load steps with commits and the Newmark terms on the diagonal"""
    rng = np.random.default_rng(8)
    for spec, sc in ((brick_block(5, 4, 3, distort=0.25, seed=3, body=(0.01, 0.0, -0.02)), 2e-3),
                     (brick_periodic_equaldof(3, 3, 2, seed=5), 2e-3), (quad_plane(7, 5, mat=J2_STEEL, distort=0.2, seed=4), 2e-3),
                     (frame2d(2, 3, 2), np.array((0.006, 0.003, 6e-5)))):
        beam = spec.groups[0].kind in (2, 3)
        tol = BEAM_RTOL if beam else RTOL
        mass = rng.uniform(0.01, 0.1, (spec.nn, spec.ndf))
        O = OracleBackend(spec, 1, soe); O.set_mass(spec.node_tags, mass)
        D = xb.DeviceModel.from_spec(spec, setup=False); D.set_mass(spec.node_tags, mass); D.setup(1, soe); D.to_device(0)
        assert D.a_size == O.a_size
        ids = O.ids()
        for s_ in range(3):
            u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * (s_ + 1); u[ids < 0] = 0
            tie(spec, u)
            O.set_trial_disp(u); D.set_trial_disp(u); D.update()
            O.apply_load(0.3 * s_); D.apply_load(0.3 * s_)
            if s_ == 2:
                for m in (O, D):
                    m.set_transient(1.0, 30.0, 5.0e3)
            A = D.form_tangent()
            assert relerr(A, O.form_tangent()) < tol
            assert np.array_equal(A, D.form_tangent())                     # entries outside the pattern stay exact zeros
            assert relerr(D.form_unbalance(), O.form_unbalance()) < tol
            O.commit(); D.commit()


@pytest.mark.parametrize("dim,loads", [(2, "uniform"), (2, "point"), (2, "both"), (3, "uniform"), (3, "point"), (3, "both"), (2, "partial"), (2, "all"), (3, "partial"), (3, "all")])
def test_beam_uniform_element_loads_device_vs_oracle(dim, loads):
    """`eleLoad -beamUniform` and `-beamPoint` (Beam2d/3dPointLoad, ForceBeamColumn2d.cpp:442,1138 / 3d.cpp:457,1314) on force beams: section forces sp inside the element iteration, fixed-end reactions p0 in
    the resisting force, both scaled by the load factor; against the oracle (pinned to ForceBeamColumn2d/3d with
    Beam2d/3dUniformLoad, tests/test_oracle.py) over a gravity ramp with commits, a step in which only the load factor
    moves, a revert to the last commit and a reset"""
    from modelspec import with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(5)
    spec = frame2d(3, 3, 2) if dim == 2 else frame3d(2, 1, 2)
    # "partial" / "all": a trapezoidal load over part of every girder (Beam2d/3dPartialUniformLoad, ForceBeamColumn2d.cpp:426-443,
    # 1073-1137; ForceBeamColumn3d.cpp:432-456, 1224-1313), alone and on top of the other two kinds
    if loads not in ("point", "partial"): spec = with_beam_gravity(spec, seed=3)
    if loads not in ("uniform", "partial"): spec = with_beam_point_loads(spec, seed=2)
    if loads in ("partial", "all"):
        from modelspec import with_beam_partial_loads
        spec = with_beam_partial_loads(spec, seed=4)
    if loads == "both":       # a second uniform load on the loaded elements: the intensities add up
        spec.beam_loads = spec.beam_loads + [(t, 0.4 * wy, 0.3 * wz, -0.5 * wa) for t, wy, wz, wa in spec.beam_loads[::2]]
    nd = 6 if dim == 2 else 12
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    sc = np.asarray((0.02, 0.02, 2e-4) if dim == 2 else (0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4))
    u = np.zeros((spec.nn, spec.ndf))

    def check():
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL

    for s_ in range(5):
        if s_ != 2:                     # step 2: the load factor moves, the displacements do not
            u = u + rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * 0.1; u[ids < 0] = 0
        lam = 0.25 * (s_ + 1)
        O.apply_load(lam); O.set_trial_disp(u)
        D.apply_load(lam); D.set_trial_disp(u); D.update()
        check()
        if s_ == 3:                     # a trial state is thrown away: back to the committed load factor and state
            O.revert(); D.revert_to_last_commit()
            check()
        else:
            O.commit(); D.commit()
    O.revert_to_start(); D.revert_to_start()
    check()
    O.apply_load(0.5); O.set_trial_disp(0.3 * u); D.apply_load(0.5); D.set_trial_disp(0.3 * u); D.update()
    check()


@pytest.mark.parametrize("numberer,soe", [(0, 0), (1, 1), (1, 2)])
def test_mixed_ndf_soil_frame_device_vs_oracle(numberer, soe):
    """BASELINE configs[4] as a real mixed-ndf Domain on the device: FourNodeQuad / J2 soil on 2-dof nodes, a forceBeamColumn
    RC frame on 3-dof nodes, `equalDOF` at the column bases (shared rows), two element batches of different dofs per node in
    one assembly -- A, B, element forces against the oracle (pinned to the live reference) over a load history with commits
    and a revert; also into BandGeneral storage"""
    from modelspec import soil_frame_2d
    rng = np.random.default_rng(11)
    spec = soil_frame_2d(nbay=2, nstory=3, ndiv=2, per_bay=4, ny=5)
    O = OracleBackend(spec, numberer, soe)
    D = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
    ids = O.ids()
    assert np.array_equal(D.ids(), ids)
    nq = len(spec.groups[0].tags)
    sc = np.array((0.02, 0.02, 2e-4))

    def check():
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e, nd in ((0, 8), (nq - 1, 8), (nq, 6), (O.ne - 1, 6)):
            assert relerr(D.element_tangent(e, nd), O.ele_tangent(e, nd)) < BEAM_RTOL
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL

    check()
    for s_ in range(4):
        u = rng.normal(0, 1.0, (spec.nn, 3)) * sc * 0.3 * (s_ + 1); u[ids < 0] = 0
        tie(spec, u)
        O.set_trial_disp(u); D.set_trial_disp(u); D.update()
        O.apply_load(0.25 * (s_ + 1)); D.apply_load(0.25 * (s_ + 1))
        check()
        if s_ == 2:
            O.revert(); D.revert_to_last_commit()
            check()
        else:
            O.commit(); D.commit()


def test_launch_and_byte_accounting():
    D = xb.DeviceModel.from_spec(brick_block(3, 3, 3), 0, 0).to_device(0)
    n0 = D.launch_count()
    D.update(); D.form_unbalance(host=False); D.form_tangent(host=False); D.synchronize()
    assert D.launch_count() - n0 == 4      # update (+ element residual), assemble_B, tangent, assemble_A
    assert D.algorithmic_bytes(2) > D.nnz * 8


def _partitioned_pass(ranks, u_global, lam):
    """update + formUnbalance + formTangent on every rank of a partition held in this process,
    with the interface exchange done by device copies (xb_exchange_local)."""
    for m in ranks:
        m.set_trial_disp(u_global[m.node_tags() - 1]); m.update(); m.apply_load(lam)
        m.form_element_resids(); m.form_element_tangents()
    xb.exchange_local(ranks, 1); xb.exchange_local(ranks, 0)
    out = []
    for m in ranks:
        B = np.empty(m.nrows); A = np.empty(m.nnz)
        m.assemble_unbalance(B); m.assemble_tangent(A)
        out.append((A, B))
    return out


@pytest.mark.parametrize("nparts,numberer,soe", [(2, 0, 0), (3, 1, 1), (8, 1, 0)])
def test_partitioned_ranks_reproduce_single_gpu_bitwise(nparts, numberer, soe):
    """Every rank's owned rows of A and B equal the single-GPU rows BIT FOR BIT: remote element
    rows are added in the same global FE_Element order, whatever the partition."""
    rng = np.random.default_rng(5)
    mk = lambda: brick_block(6, 5, 7, mat=J2_STEEL, distort=0.2, seed=9, body=(0.0, 0.01, -0.02))
    spec = mk()
    G = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
    O = OracleBackend(spec, numberer, soe)
    gptr, _ = G.pattern()
    ranks = [xb.DeviceModel.from_spec(mk(), numberer, soe, nparts, r).to_device(0) for r in range(nparts)]
    ids = G.ids()
    for s in range(3):
        u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 3)); u[ids < 0] = 0
        lam = 0.4 * s
        G.set_trial_disp(u); G.update(); G.apply_load(lam)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        O.set_trial_disp(u); O.apply_load(lam)
        assert relerr(Ag, O.form_tangent()) < RTOL and relerr(Bg, O.form_unbalance()) < RTOL
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, lam)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            ptr, _ = m.pattern()
            for lr, q in enumerate(rows):
                assert np.array_equal(A[ptr[lr]:ptr[lr + 1]], Ag[gptr[q]:gptr[q + 1]])
        G.commit(); O.commit()
        for m in ranks:
            m.commit()


@pytest.mark.parametrize("shape,nparts", [("brick", 3), ("soilcolumn", 2), ("frame2d", 4)])
def test_partitioned_equal_dof_reproduces_single_gpu_bitwise(shape, nparts):
    """`equalDOF` across ranks: the tie group's rank assembles the shared rows from its own element rows and the ones
    it receives, in global (FE_Element, element dof) order -- the same bits as the single-GPU run."""
    rng = np.random.default_rng(8)
    if shape == "brick":
        mk, sc = (lambda: brick_periodic_equaldof(6, 4, 3, seed=41)), 2e-3
    elif shape == "soilcolumn":
        mk, sc = (lambda: soil_column_equaldof(24, seed=42)), 2e-3
    else:
        mk, sc = (lambda: frame2d_diaphragm_equaldof(4, 3, 2)), np.array((0.006, 0.003, 6e-5))
    spec = mk()
    mass = rng.uniform(0.01, 0.1, (spec.nn, spec.ndf))

    def build(*part):
        m = xb.DeviceModel.from_spec(mk(), setup=False); m.set_mass(spec.node_tags, mass); m.setup(1, 0, *part)
        return m.to_device(0)
    G = build()
    gptr, _ = G.pattern()
    ranks = [build(nparts, r) for r in range(nparts)]
    assert sorted(np.concatenate([m.row_eqns() for m in ranks]).tolist()) == list(range(G.neq))
    ids = G.ids()
    for s in range(3):
        u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * (s + 1); u[ids < 0] = 0
        tie(spec, u)
        if s == 2:     # transient terms: nodal masses of every dof on a shared equation
            for m in [G] + ranks:
                m.set_transient(1.0, 40.0, 8.0e3)
        lam = 0.4 * s
        G.set_trial_disp(u); G.update(); G.apply_load(lam)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, lam)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            ptr, _ = m.pattern()
            for lr, q in enumerate(rows):
                assert np.array_equal(A[ptr[lr]:ptr[lr + 1]], Ag[gptr[q]:gptr[q + 1]])
        G.commit()
        for m in ranks:
            m.commit()


def test_partitioned_quad_and_scattered_partition():
    rng = np.random.default_rng(6)
    mk = lambda: quad_plane(12, 9, mat=J2_STEEL, lx=12.0, ly=9.0, distort=0.2, seed=2)
    spec = mk()
    part = rng.integers(0, 3, spec.ne).astype(np.int32)          # worst case: no locality at all
    G = xb.DeviceModel.from_spec(spec, 1, 1).to_device(0)
    gptr, _ = G.pattern()
    ranks = [xb.DeviceModel.from_spec(mk(), 1, 1, 3, r, part).to_device(0) for r in range(3)]
    u = rng.normal(0, 4e-3, (spec.nn, 2)); u[G.ids() < 0] = 0
    G.set_trial_disp(u); G.update(); G.apply_load(1.0)
    Ag, Bg = G.form_tangent(), G.form_unbalance()
    for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, 1.0)):
        rows = m.row_eqns(); ptr, _ = m.pattern()
        assert np.array_equal(B, Bg[rows])
        assert np.array_equal(A, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows]))


@pytest.mark.skipif(not have_metis(), reason="oracle/_ref/libmetis_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("nparts", [4, 8])
def test_soil_structure_metis_partition_reproduces_single_gpu(nparts):
    """BASELINE configs[4] in small: a mixed mesh (J2 soil + elastic footing/pier: two element batches whose
    FE order interleaves) split by METIS_PartGraphKway exactly as the reference's DomainPartitioner would
    (reference's own METIS 4 on Domain::buildEleGraph's element graph).  The single-GPU result matches the
    oracle; every rank's owned rows match the single-GPU rows bit for bit."""
    rng = np.random.default_rng(8)
    mk = lambda: soil_structure_block(8, 8, 6, distort=0.15, seed=4)
    spec = mk()
    part = metis_partition(spec, nparts)
    assert np.bincount(part, minlength=nparts).min() > 0
    G = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    O = OracleBackend(spec, 1, 0)
    gptr, _ = G.pattern()
    ranks = [xb.DeviceModel.from_spec(mk(), 1, 0, nparts, r, part).to_device(0) for r in range(nparts)]
    assert np.array_equal(ranks[0].partition(G.ne), part)
    ids = G.ids()
    for s in range(2):
        u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 3)); u[ids < 0] = 0
        lam = 0.5 * (s + 1)
        G.set_trial_disp(u); G.update(); G.apply_load(lam)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        O.set_trial_disp(u); O.apply_load(lam)
        assert relerr(Ag, O.form_tangent()) < RTOL and relerr(Bg, O.form_unbalance()) < RTOL
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, lam)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            assert np.array_equal(A, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows]))
        G.commit(); O.commit()
        for m in ranks:
            m.commit()


NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import xara_b200 as xb
from modelspec import J2_STEEL, brick_block
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
spec = brick_block(8, 6, 10, mat=J2_STEEL, distort=0.2, seed=3)
m = xb.DeviceModel.from_spec(spec, 1, 0, world, rank).to_device(rank)
box = [xb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, 0)
m.comm_init(box[0])
rng = np.random.default_rng(1)
u = rng.normal(0, 3e-3, (spec.nn, 3))
m.set_trial_disp(u[m.node_tags() - 1]); m.update(); m.apply_load(0.7)
A, B = m.form_tangent(), m.form_unbalance()          # NCCL exchange inside
np.savez(sys.argv[1] + f".{{rank}}.npz", A=A, B=B, rows=m.row_eqns(), ptr=m.pattern()[0])
dist.barrier()
print("rank", rank, "done")
'''


@pytest.mark.skipif(xb.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_nccl_exchange_matches_single_gpu(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(NCCL_WORKER.format(root=root))
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    spec = brick_block(8, 6, 10, mat=J2_STEEL, distort=0.2, seed=3)
    G = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    u = np.random.default_rng(1).normal(0, 3e-3, (spec.nn, 3))
    G.set_trial_disp(u); G.update(); G.apply_load(0.7)
    Ag, Bg = G.form_tangent(), G.form_unbalance()
    gptr, _ = G.pattern()
    for rank in range(2):
        d = np.load(out + f".{rank}.npz")
        assert np.array_equal(d["B"], Bg[d["rows"]])
        assert np.array_equal(d["A"], np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in d["rows"]]))


def test_frame_fibre_beams_vs_oracle_history():
    """forceBeamColumn + FiberSection2d (Steel02 / Concrete02): cyclic sway history with commits and a
    revertToLastCommit, device against the oracle"""
    spec = frame2d(3, 4, 2)
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL          # initial (elastic) stiffness
    y = spec.crd[:, 1] / spec.crd[:, 1].max()
    rng = np.random.default_rng(3)
    amp = [0.4, 1.2, 2.5, 1.0, -1.5, -3.0, 0.5, 3.5]                        # inches of roof drift, cycling
    for s, a in enumerate(amp):
        u = np.zeros((spec.nn, 3))
        u[:, 0] = a * y ** 1.5; u[:, 1] = -0.01 * y; u[:, 2] = -1.5 * a * y ** 0.5 / spec.crd[:, 1].max()
        u += rng.normal(0, 1.0, u.shape) * (2e-3, 1e-3, 2e-5)
        u[ids < 0] = 0
        assert O.set_trial_disp(u) == 0
        D.set_trial_disp(u); D.update(); D.apply_load(0.1 * s); O.apply_load(0.1 * s)
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        if s == 4:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
            assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); D.commit()
    # the history went well past yield: the tangent is far from the initial one
    D0 = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    assert relerr(D.form_tangent(), D0.form_tangent()) > 0.05


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("pdelta", [0, 1])
def test_joint_offsets_device_vs_oracle(pdelta, dim):
    """`geomTransf Linear | PDelta ... -jntOffset` on 2D / 3D force beams (rigid end zones): sway history under gravity
    element loads with commits, a revert and a reset, with Rayleigh damping terms in a transient step at the end; device
    against the oracle (pinned to Linear / PDeltaCrdTransf2d / 3d with offsets)"""
    from modelspec import with_joint_offsets, with_beam_gravity, with_pdelta
    rng = np.random.default_rng(21)
    base = frame2d(3, 2, 2, gravity=-80.0) if dim == 2 else frame3d(1, 2, 2, gravity=-40.0)
    spec = with_joint_offsets(with_beam_gravity(base, w=-0.08 if dim == 2 else -0.06, seed=1), seed=3)
    if pdelta: spec = with_pdelta(spec)
    ndf, nd = spec.ndf, 2 * spec.ndf
    mass = np.zeros((spec.nn, ndf)); mass[:, :dim] = 0.05
    O = OracleBackend(spec, 1, 0); O.set_mass(spec.node_tags, mass)
    D = xb.DeviceModel.from_spec(spec, setup=False); D.set_mass(spec.node_tags, mass); D.setup(1, 0); D.to_device(0)
    ids = O.ids()
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
    hc = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hc.max(); h = hc / H
    pattern = rng.normal(0, 1.0, (spec.nn, ndf)) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))

    def check():
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL and relerr(D.element_tangent(e, nd), O.ele_tangent(e, nd)) < BEAM_RTOL
    for s, a in enumerate([0.2, 0.5, 0.8, 1.1, 1.4] if dim == 2 else [0.2, 0.4, 0.6, 0.8, 1.0]):
        u = np.zeros((spec.nn, ndf)); u[:, 0] = a * h ** 1.5; u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        if dim == 3: u[:, 1] = 0.5 * a * h ** 1.5; u[:, 3] = 0.7 * a * h ** 0.5 / H
        u += pattern * (a / 0.5); u[ids < 0] = 0
        O.apply_load(0.2 * (s + 1)); assert O.set_trial_disp(u) == 0
        D.apply_load(0.2 * (s + 1)); D.set_trial_disp(u); D.update()
        check()
        if s == 4:
            O.revert(); D.revert_to_last_commit(); check()
        else:
            O.commit(); D.commit()
    # a transient step with all four Rayleigh factors: the damping forces go through the offsets as well
    for m in (O, D):
        m.set_rayleigh(0.3, 0.002, 0.001, 0.0015); m.set_transient(1.0, 50.0, 5000.0)
    v = rng.normal(0, 0.5, (spec.nn, ndf)); acc = rng.normal(0, 5.0, (spec.nn, ndf)); v[ids < 0] = 0; acc[ids < 0] = 0
    O.set_vel_accel(v, acc); D.set_vel_accel(v, acc)
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL


@pytest.mark.parametrize("loads", [0, 1])
def test_corotational_device_vs_oracle(loads):
    """`geomTransf Corotational` on 2D force beams (CorotCrdTransf2d): large-displacement sway history (drifts of several
    per cent), with and without element loads, commits, a revert to the last commit and a reset; a Newmark step with
    mass-proportional damping at the end (stiffness-proportional Rayleigh terms on corotational beams are refused).
    Device against the oracle (pinned to the reference's classes, tests/test_oracle.py)."""
    from modelspec import with_corot, with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(33)
    spec = frame2d(3, 2, 2, gravity=-80.0)
    if loads: spec = with_beam_point_loads(with_beam_gravity(spec, w=-0.08, seed=1), P=-2.0, seed=2)
    spec = with_corot(spec)
    mass = np.zeros((spec.nn, 3)); mass[:, :2] = 0.05
    O = OracleBackend(spec, 1, 0); O.set_mass(spec.node_tags, mass)
    D = xb.DeviceModel.from_spec(spec, setup=False); D.set_mass(spec.node_tags, mass); D.setup(1, 0); D.to_device(0)
    ids = O.ids()
    H = spec.crd[:, 1].max(); h = spec.crd[:, 1] / H
    pattern = rng.normal(0, 1.0, (spec.nn, 3)) * (2e-3, 1e-3, 2e-5)

    def check():
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_resid(e, 6), O.ele_resid(e, 6)) < BEAM_RTOL and relerr(D.element_tangent(e, 6), O.ele_tangent(e, 6)) < BEAM_RTOL
    check()
    for rep in range(2):
        # (with element loads the force-based iteration gives up earlier on the way back: a milder history)
        for s, a in enumerate([0.5, 1.0, 1.5, 2.5, 3.5, 3.0] if loads else [0.5, 1.5, 3.0, 5.0, 7.0, 5.5]):
            u = np.zeros((spec.nn, 3)); u[:, 0] = a * h ** 1.5; u[:, 2] = -1.5 * a * h ** 0.5 / H
            u += pattern * (a / 0.5); u[ids < 0] = 0
            O.apply_load(0.2 * (s + 1)); assert O.set_trial_disp(u) == 0
            D.apply_load(0.2 * (s + 1)); D.set_trial_disp(u); D.update()
            check()
            if s == 4:
                O.revert(); D.revert_to_last_commit(); check()
            else:
                O.commit(); D.commit()
        if rep == 0:      # `reset`, then the same history again: same numbers as the first time
            A1 = D.form_tangent().copy()
            O.revert_to_start(); D.revert_to_start(); check()
        else:
            assert np.array_equal(A1, D.form_tangent())
    # far from the linear transformation's tangent
    Dl = xb.DeviceModel.from_spec(frame2d(3, 2, 2, gravity=-80.0), 1, 0).to_device(0)
    assert relerr(D.form_tangent(), Dl.form_tangent()) > 0.05
    with pytest.raises(Exception):
        D.set_rayleigh(0.3, 0.002, 0.0, 0.0)
    for m in (O, D):
        m.set_rayleigh(0.3, 0.0, 0.0, 0.0); m.set_transient(1.0, 50.0, 5000.0)
    v = rng.normal(0, 0.5, (spec.nn, 3)); acc = rng.normal(0, 5.0, (spec.nn, 3)); v[ids < 0] = 0; acc[ids < 0] = 0
    O.set_vel_accel(v, acc); D.set_vel_accel(v, acc)
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL


@pytest.mark.parametrize("dim", [2, 3])
def test_gravity_then_pushover_load_const_device_vs_oracle(dim):
    """Gravity (nodal loads, `eleLoad -beamUniform` / `-beamPoint`) ramped to its full value, xb_load_const +
    xb_apply_load(0) (`loadConst -time 0`), a lateral pattern through xb_set_nodal_loads, then the push: the frozen loads --
    nodal and element -- keep their factor, the new pattern follows the domain time; a revert to the last commit on the way.
    Device against the oracle (pinned to the reference for this sequence, tests/test_oracle.py)."""
    from modelspec import with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(13)
    mk = (lambda: frame2d(2, 2, 2, lateral=0.0, gravity=-60.0)) if dim == 2 else (lambda: frame3d(1, 1, 2, lateral=(0.0, 0.0), gravity=-30.0))
    spec = with_beam_point_loads(with_beam_gravity(mk(), w=-0.08, seed=1), P=-2.0, seed=2)
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    pattern = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))

    def step(a, lam, commit=True):
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5; u[:, 1 if dim == 2 else 2] = -0.01 * h
        u += pattern * (a / 0.06); u[ids < 0] = 0
        O.apply_load(lam); assert O.set_trial_disp(u) == 0
        D.apply_load(lam); D.set_trial_disp(u); D.update()
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL
        if commit: O.commit(); D.commit()
    for s in range(3):
        step(0.02 * (s + 1), (s + 1) / 3.0)
    O.load_const(); O.apply_load(0.0); D.load_const(); D.apply_load(0.0)
    top = [int(t) for t, c in zip(spec.node_tags, spec.crd) if c[1 if dim == 2 else 2] == H]
    for t in top:
        v = np.zeros(spec.ndf); v[0] = 12.0
        O.add_load(t, v); D.add_load(t, v)
    step(0.07, 0.0)
    step(0.3, 0.3)
    step(0.6, 0.6, commit=False)
    O.revert(); D.revert_to_last_commit()             # back to time 0.3: the frozen loads stay in full
    assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL and relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
    step(0.6, 0.6)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (the rules come from the reference's BeamIntegration classes)")
@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("dim", [2, 3])
def test_beam_integration_rules_device_vs_oracle(dim, kind):
    """forceBeamColumn -integration Legendre | Radau | NewtonCotes | Trapezoidal (xb_set_beam_integration: the section
    locations and weights the reference's BeamIntegration object returns, per element), with uniform and point element
    loads, commits, a revert and a reset: device against the oracle (pinned to the reference's classes for every rule)"""
    from modelspec import with_beam_integration, with_beam_gravity, with_beam_point_loads
    rng = np.random.default_rng(11)
    spec = frame2d(2, 2, 2, nip=4) if dim == 2 else frame3d(1, 1, 2, nip=5)
    spec = with_beam_integration(with_beam_point_loads(with_beam_gravity(spec, seed=3), seed=2), kind)
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
    for s, a in enumerate([0.2, 0.5, 0.8, 1.1]):        # (the history of the CPU test against the reference)
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5
        u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        u += rng.normal(0, 1.0, u.shape) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))
        u[ids < 0] = 0
        O.apply_load(0.25 * (s + 1)); assert O.set_trial_disp(u) == 0
        D.apply_load(0.25 * (s + 1)); D.set_trial_disp(u); D.update()
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, O.ne // 2, O.ne - 1):
            assert relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL
        if s == 3:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); D.commit()
    O.revert_to_start(); D.revert_to_start()
    O.apply_load(0.0); D.apply_load(0.0)
    z = np.zeros((spec.nn, spec.ndf)); O.set_trial_disp(z); D.set_trial_disp(z); D.update()
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL


@pytest.mark.parametrize("bars", ["steel01", "elasticpp"])
@pytest.mark.parametrize("dim", [2, 3])
def test_steel01_elastic_fibres_device_vs_oracle(dim, bars):
    """Steel01 and Elastic fibres inside FiberSection2d / FiberSection3d (the new uniaxial kinds as ordinary fibres):
    sway history with commits and a revert, device against the oracle"""
    from modelspec import steel01_elastic_frame
    rng = np.random.default_rng(8)
    spec = steel01_elastic_frame(dim, bars)          # "elasticpp": ElasticPPMaterial bars
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    A0 = O.form_tangent().copy()
    for s, a in enumerate([0.3, 0.8, 1.4, 2.0, 2.4]):
        u = np.zeros((spec.nn, spec.ndf))
        u[:, 0] = a * h ** 1.5
        u[:, 2 if dim == 2 else 4] = -1.5 * a * h ** 0.5 / H
        u += rng.normal(0, 1.0, u.shape) * ((2e-3, 1e-3, 2e-5) if dim == 2 else (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5))
        u[ids < 0] = 0
        O.apply_load(0.2 * (s + 1)); assert O.set_trial_disp(u) == 0
        D.apply_load(0.2 * (s + 1)); D.set_trial_disp(u); D.update()
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        if s == 4:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); D.commit()
    assert relerr(O.form_tangent(), A0) > 0.05


@pytest.mark.parametrize("dim", [2, 3])
def test_pdelta_transformation_device_vs_oracle(dim):
    """forceBeamColumn under `geomTransf PDelta` (PDeltaCrdTransf2d.cpp / PDeltaCrdTransf3d.cpp: geometric stiffness N/L,
    leaning-column shear): sway history under gravity with commits and a revert to the last commit, device against the
    oracle (pinned to the reference's classes, incl. the 3D element's stale relative displacements after a revert)"""
    from modelspec import with_beam_gravity, with_pdelta
    rng = np.random.default_rng(6)
    mk = (lambda: frame2d(2, 3, 2, gravity=-120.0)) if dim == 2 else (lambda: frame3d(1, 1, 2, gravity=-40.0))
    spec = with_pdelta(with_beam_gravity(mk(), seed=4))
    O = OracleBackend(spec, 1, 0); Ol = OracleBackend(with_beam_gravity(mk(), seed=4), 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    hcol = spec.crd[:, 1] if dim == 2 else spec.crd[:, 2]
    H = hcol.max(); h = hcol / H
    nd = 6 if dim == 2 else 12
    differs = False
    for s, a in enumerate([0.3, 0.7, 1.1, 1.5, 1.9, 2.3]):
        u = np.zeros((spec.nn, spec.ndf))
        if dim == 2:
            u[:, 0] = a * h ** 1.5; u[:, 1] = -0.01 * h; u[:, 2] = -1.5 * a * h ** 0.5 / H
            u += rng.normal(0, 1.0, u.shape) * (2e-3, 1e-3, 2e-5)
        else:
            u[:, 0] = a * h ** 1.5; u[:, 1] = 0.6 * a * h ** 1.5; u[:, 2] = -0.01 * h
            u[:, 3] = 0.9 * a * h ** 0.5 / H; u[:, 4] = -1.5 * a * h ** 0.5 / H
            u += rng.normal(0, 1.0, u.shape) * (2e-3, 2e-3, 1e-3, 2e-5, 2e-5, 2e-5)
        u[ids < 0] = 0
        lam = 0.2 * (s + 1)
        for m in (O, Ol):
            m.apply_load(lam); assert m.set_trial_disp(u) == 0
        D.apply_load(lam); D.set_trial_disp(u); D.update()
        Ao, Bo = O.form_tangent(), O.form_unbalance()
        assert relerr(D.form_tangent(), Ao) < BEAM_RTOL and relerr(D.form_unbalance(), Bo) < BEAM_RTOL
        for e in range(O.ne):
            assert relerr(D.element_tangent(e, nd), O.ele_tangent(e, nd)) < BEAM_RTOL and relerr(D.element_resid(e, nd), O.ele_resid(e, nd)) < BEAM_RTOL
        differs = differs or (relerr(Ao, Ol.form_tangent()) > 1e-8 and relerr(Bo, Ol.form_unbalance()) > 1e-8)
        if s == 5:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); Ol.commit(); D.commit()
    assert differs


@pytest.mark.parametrize("ndiv", [1, 4])
def test_aggregator_cantilever_device_vs_oracle(ndiv):
    """forceBeamColumn over `section Aggregator` (Elastic on P, Steel01 on Mz -- the section BASELINE configs[0]'s script
    defines; SectionAggregator.cpp:316-500, Steel01.cpp:68-196): pushed past yield, released, reversed, with commits, a
    revertToLastCommit and a reset, device against the oracle (which is pinned to the reference's classes)"""
    from modelspec import cantilever2d_aggregator
    spec = cantilever2d_aggregator(ndiv=ndiv)
    O = OracleBackend(spec, 0, 0)
    D = xb.DeviceModel.from_spec(spec, 0, 0).to_device(0)
    ids = O.ids()
    A0 = O.form_tangent()
    assert relerr(D.form_tangent(), A0) < BEAM_RTOL
    rng = np.random.default_rng(9)
    amp = np.array([10.0, 0.005, 0.04])
    yielded = False
    for s, f in enumerate([0.1, 0.5, 1.0, 0.7, 0.2, -0.5, -1.0, 0.3]):
        h = np.linspace(0.0, 1.0, spec.nn)[:, None]
        u = f * amp * h ** 2 + rng.normal(0, 1.0, (spec.nn, 3)) * amp * 0.01; u[ids < 0] = 0
        assert O.set_trial_disp(u) == 0
        D.set_trial_disp(u); D.update(); D.apply_load(0.1 * s); O.apply_load(0.1 * s)
        Ao = O.form_tangent()
        assert relerr(D.form_tangent(), Ao) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in range(O.ne):
            assert relerr(D.element_tangent(e, 6), O.ele_tangent(e, 6)) < BEAM_RTOL and relerr(D.element_resid(e, 6), O.ele_resid(e, 6)) < BEAM_RTOL
        yielded = yielded or relerr(Ao, A0) > 1e-3
        if s == 4:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL and relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); D.commit()
    assert yielded
    O.revert_to_start(); D.revert_to_start()
    z = np.zeros((spec.nn, 3))
    O.set_trial_disp(z); D.set_trial_disp(z); D.update()
    assert relerr(D.form_tangent(), A0) < BEAM_RTOL


def test_ex2b_as_written_device_vs_golden():
    """BASELINE configs[0] as the script is written (gravity under LoadControl, loadConst -time 0 = xb_load_const +
    xb_apply_load, the lateral pattern = xb_set_nodal_loads, pushover under DisplacementControl), driven through the
    C-ABI, against the history the UNMODIFIED reference produced with `system BandGeneral` (tests/golden/ex2b_as_written.npz)"""
    from modelspec import cantilever2d_aggregator, ex2b_drive
    g = np.load(os.path.join(GOLD, "ex2b_as_written.npz"))
    spec = cantilever2d_aggregator(ndiv=1, H=0.0, V=-2000.0)
    D = xb.DeviceModel.from_spec(spec, 0, 0).to_device(0)
    ids = D.ids(); ptr, idx = D.pattern(); neq = D.neq

    def solve(A, b):
        M = np.zeros((neq, neq))
        for c in range(neq): M[idx[ptr[c]:ptr[c + 1]], c] = A[ptr[c]:ptr[c + 1]]
        return np.linalg.solve(M, b)

    grav_u, lam, u = ex2b_drive(D, solve, ids, True)
    assert relerr(grav_u, g["grav_u"]) < 1e-10
    assert relerr(lam, g["push_lam"]) < 1e-9 and relerr(u, g["push_u"]) < 1e-9


@pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_gravity_then_pushover_on_device_path():
    """An RC frame the way it is usually analysed: gravity with `eleLoad -beamUniform` / `-beamPoint` and nodal loads in five
    LoadControl steps, `loadConst -time 0`, a lateral `pattern Plain`, a second `analysis Static` that pushes in eight steps --
    on the reference's own objects, unmodified (CPU) and with both analyses' integrators routed to the device (the second
    one re-attaches to the device model of the first: xb_load_const freezes nodal and element loads, xb_set_nodal_loads
    brings the push pattern).  Same iteration counts in both phases, same displacements."""
    from modelspec import GLUE_SO, RefBackend, with_beam_gravity, with_beam_point_loads

    def run(tol, glue):
        spec = with_beam_point_loads(with_beam_gravity(frame2d(2, 3, 2, lateral=0.0, gravity=-60.0), w=-0.08, seed=1), P=-2.0, seed=2)
        R = RefBackend(spec, defer_setup=True, so=GLUE_SO if glue else None)
        (R.setup_glue_loadcontrol if glue else R.setup_loadcontrol)(1, 0, 0.2, test=0, tol=tol, max_iter=25)
        rc, it1, nm1 = R.analyze_static(5)
        assert rc == 0
        R.load_const(0.0)
        H = spec.crd[:, 1].max()
        for t, c in zip(spec.node_tags, spec.crd):
            if c[1] > 0 and c[0] == 0.0: R.add_load(int(t), (22.0 * c[1] / H, 0.0, 0.0))
        (R.setup_glue_loadcontrol if glue else R.setup_loadcontrol)(1, 0, 0.125, test=0, tol=tol, max_iter=25)
        rc, it2, nm2 = R.analyze_static(8)
        assert rc == 0
        u = R.glue_trial_disp() if glue else R.get_trial_disp()
        return np.concatenate([it1, it2]), np.vstack([nm1, nm2]), u, (R.glue_counts() if glue else None)

    best = None
    for tol in (1e-6, 1e-7, 1e-8, 1e-9):
        it, nm, u, _ = run(tol, False)
        margin = min(min(nm[s, it[s] - 2] / tol if it[s] > 1 else 1e9, tol / max(nm[s, it[s] - 1], 1e-300)) for s in range(len(it)))
        if best is None or margin > best[0]:
            best = (margin, tol, it, nm, u)
    margin, tol, it_cpu, nm_cpu, u_cpu = best
    assert margin >= 2.0 and it_cpu[5:].max() >= 4, (margin, it_cpu)
    it_dev, nm_dev, u_dev, (calls, launches) = run(tol, True)
    assert it_dev.tolist() == it_cpu.tolist()
    for s in range(len(it_cpu)):
        assert np.allclose(nm_dev[s, :it_dev[s] - 1], nm_cpu[s, :it_cpu[s] - 1], rtol=1e-5, atol=1e-11)
    assert relerr(u_dev, u_cpu) < 1e-8
    assert calls[3] == 8 and launches > 0


@pytest.mark.skipif(not have_glue(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_runs_ex2b_as_written_on_device_path():
    """tests/Ex2b.Canti2D.InelasticSection.Push.py statement by statement on the reference's OWN objects (Domain,
    SectionAggregator, Steel01, ElasticMaterial, ForceBeamColumn2d, PlainHandler, PlainNumberer, BandGenLinSOE,
    NewtonRaphson, CTestNormDispIncr / CTestEnergyIncr, LoadControl, DisplacementControl) with the integrators'
    model-facing calls routed to the device (oracle/ref_glue.cpp): the device model outlives the first analysis,
    `loadConst` and the pushover pattern reach it through xb_load_const / xb_set_nodal_loads, A lands in the
    BandGenLinSOE's own array.  Iteration counts, load factors and displacements are those of the unmodified run."""
    from modelspec import ex2b_as_written
    g = np.load(os.path.join(GOLD, "ex2b_as_written.npz"))
    d = ex2b_as_written(glue=True)
    assert d["grav_iters"].tolist() == g["grav_iters"].tolist() and d["push_iters"].tolist() == g["push_iters"].tolist()
    assert relerr(d["grav_u"], g["grav_u"]) < 1e-10
    assert relerr(d["push_lam"], g["push_lam"]) < 1e-9 and relerr(d["push_u"], g["push_u"]) < 1e-9
    assert d["calls"][3] == 50 and d["launches"] > 0


def test_frame3d_fibre_beams_vs_oracle_history():
    """forceBeamColumn in 3D (ForceBeamColumn3d) + FiberSection3d (Steel02 / Concrete02, elastic torsion):
    cyclic biaxial sway + twist history with commits and a revertToLastCommit, device against the oracle"""
    spec = frame3d(2, 2, 3, ndiv=2)
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL          # initial (elastic) stiffness
    H = spec.crd[:, 2].max()
    z = spec.crd[:, 2] / H
    xc, yc = spec.crd[:, 0] - spec.crd[:, 0].mean(), spec.crd[:, 1] - spec.crd[:, 1].mean()
    rng = np.random.default_rng(5)
    amp = [0.2, 0.6, 1.2, 0.5, -0.7, -1.4, 0.3, 0.6]                        # inches of roof drift, cycling
    for s, a in enumerate(amp):
        th = 2e-3 * a * z                                                   # floor twist about the vertical axis
        u = np.zeros((spec.nn, 6))
        u[:, 0] = a * z ** 1.5 - th * yc; u[:, 1] = 0.6 * a * z ** 1.5 + th * xc; u[:, 2] = -0.01 * z
        u[:, 3] = -0.6 * 1.5 * a * z ** 0.5 / H; u[:, 4] = 1.5 * a * z ** 0.5 / H; u[:, 5] = th
        u += rng.normal(0, 1.0, u.shape) * (2e-3, 2e-3, 5e-4, 1e-5, 1e-5, 1e-5)
        u[ids < 0] = 0
        assert O.set_trial_disp(u) == 0
        D.set_trial_disp(u); D.update(); D.apply_load(0.1 * s); O.apply_load(0.1 * s)
        assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
        assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        for e in (0, spec.groups[0].conn.shape[0] - 1):
            assert relerr(D.element_tangent(e, 12), O.ele_tangent(e, 12)) < BEAM_RTOL
            assert relerr(D.element_resid(e, 12), O.ele_resid(e, 12)) < BEAM_RTOL
        if s == 4:
            O.revert(); D.revert_to_last_commit()
            assert relerr(D.form_tangent(), O.form_tangent()) < BEAM_RTOL
            assert relerr(D.form_unbalance(), O.form_unbalance()) < BEAM_RTOL
        else:
            O.commit(); D.commit()
    D0 = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    assert relerr(D.form_tangent(), D0.form_tangent()) > 0.05              # well past yield


def test_partitioned_frame3d_matches_single_gpu():
    spec_fn = lambda: frame3d(3, 2, 3, ndiv=2)
    spec = spec_fn()
    G = xb.DeviceModel.from_spec(spec, 1, 1).to_device(0)
    gptr, _ = G.pattern()
    ranks = [xb.DeviceModel.from_spec(spec_fn(), 1, 1, 3, r).to_device(0) for r in range(3)]
    H = spec.crd[:, 2].max()
    z = spec.crd[:, 2] / H
    for a in (0.6, 1.8):
        u = np.zeros((spec.nn, 6)); u[:, 0] = a * z ** 1.5; u[:, 1] = 0.5 * a * z ** 1.5
        u[:, 3] = -0.5 * 1.5 * a * z ** 0.5 / H; u[:, 4] = 1.5 * a * z ** 0.5 / H
        u[G.ids() < 0] = 0
        G.set_trial_disp(u); G.update(); G.apply_load(1.0)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, 1.0)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            assert np.array_equal(A, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows]))
        G.commit()
        for m in ranks:
            m.commit()


def test_partitioned_frame_matches_single_gpu():
    spec_fn = lambda: frame2d(4, 5, 2)
    spec = spec_fn()
    G = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    gptr, _ = G.pattern()
    ranks = [xb.DeviceModel.from_spec(spec_fn(), 1, 0, 3, r).to_device(0) for r in range(3)]
    y = spec.crd[:, 1] / spec.crd[:, 1].max()
    for a in (0.8, 2.4):
        u = np.zeros((spec.nn, 3)); u[:, 0] = a * y ** 1.5; u[:, 2] = -1.5 * a * y ** 0.5 / spec.crd[:, 1].max()
        u[G.ids() < 0] = 0
        G.set_trial_disp(u); G.update(); G.apply_load(1.0)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, 1.0)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            assert np.array_equal(A, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows]))
        G.commit()
        for m in ranks:
            m.commit()


@pytest.mark.parametrize("name", ["newmark_brick_j2", "newmark_frame2d", "newmark_frame3d", "rayleigh_brick_j2",
                                  "rayleigh_quad_j2", "rayleigh_frame2d", "rayleigh_frame3d",
                                  "rayleigh_soilcolumn_equaldof", "rayleigh_frame2d_equaldof", "rayleigh_quad_planestress",
                                  "rayleigh_quad_planestress_j2", "rayleigh_frame2d_pdelta", "rayleigh_frame3d_pdelta",
                                  "rayleigh_frame2d_rho", "rayleigh_frame3d_rho"])
def test_newmark_device_vs_golden_reference_history(name):
    """Newmark (displacement form, nodal masses): the device replays the history recorded from the
    reference's own Newmark integrator -- c1 K + c3 M tangent, P - M a - R unbalance, predictor,
    response update -- and matches A, B, velocities and accelerations.  The rayleigh_* histories add
    `rayleigh alphaM betaK betaKinit betaKcomm` and element masses from the material density: element tangent
    c1 Kt + c2 (alphaM M + betaK Kt + betaK0 K0 + betaKc Kc) + c3 M, residual getResistingForceIncInertia."""
    from golden_cases import RAYLEIGH_CASES, TRANSIENT_CASES
    from test_oracle import drive_transient_vs_golden
    mk, *_ = {**TRANSIENT_CASES, **RAYLEIGH_CASES}[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = mk()
    D = xb.DeviceModel.from_spec(spec, setup=False)
    D.set_mass(spec.node_tags, g["mass"])
    D.setup(1, 1); D.to_device(0)
    assert np.array_equal(D.ids(), g["ids"])
    tol = BEAM_RTOL if spec.groups[0].kind in (2, 3) else 1e-11

    def check(A, Ag, B, Bg, bscale):
        assert relerr(A, Ag) < tol
        assert np.abs(B - Bg).max() <= tol * bscale

    drive_transient_vs_golden(D, g, name, check)


def test_partitioned_rayleigh_transient_matches_single_gpu():
    """damping and element-mass terms on a partitioned model: interface rows (tangent, mass, damping forces)
    travel like the static ones, owned rows are bitwise those of the single-GPU run"""
    from golden_cases import J2_STEEL_RHO, RAYLEIGH, newmark_coeffs
    mk = lambda: brick_block(5, 4, 6, mat=J2_STEEL_RHO, distort=0.2, seed=11)
    spec = mk()
    (c1, c2, c3), _ = newmark_coeffs(0.5, 0.25, 0.02)
    rng = np.random.default_rng(12)

    def prep(D):
        D.to_device(0); D.set_rayleigh(*RAYLEIGH); D.set_transient(c1, c2, c3)
        return D

    G = prep(xb.DeviceModel.from_spec(spec, 1, 0))
    O = OracleBackend(spec, 1, 0); O.set_rayleigh(*RAYLEIGH); O.set_transient(c1, c2, c3)
    gptr, _ = G.pattern()
    ranks = [prep(xb.DeviceModel.from_spec(mk(), 1, 0, 3, r)) for r in range(3)]
    ids = G.ids()
    for s in range(2):
        u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 3)); u[ids < 0] = 0
        v = rng.normal(0, 0.05, (spec.nn, 3)); v[ids < 0] = 0
        a = rng.normal(0, 2.0, (spec.nn, 3)); a[ids < 0] = 0
        G.set_trial_disp(u); G.set_vel_accel(v, a); G.update(); G.apply_load(0.5)
        Ag, Bg = G.form_tangent(), G.form_unbalance()
        O.set_trial_disp(u); O.set_vel_accel(v, a); O.apply_load(0.5)
        assert relerr(Ag, O.form_tangent()) < 1e-11 and relerr(Bg, O.form_unbalance()) < 1e-11
        for m in ranks:
            m.set_vel_accel(v[m.node_tags() - 1], a[m.node_tags() - 1])
        for m, (A, B) in zip(ranks, _partitioned_pass(ranks, u, 0.5)):
            rows = m.row_eqns()
            assert np.array_equal(B, Bg[rows])
            assert np.array_equal(A, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows]))
        G.commit(); O.commit()
        for m in ranks:
            m.commit()


def test_newmark_time_history_counts_match_oracle():
    """a transient run (Newmark average acceleration, nodal masses, alphaM damping) driven by the
    oracle and by the device: same Newton iteration counts per step, same response"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from golden_cases import newmark_coeffs
    spec = brick_block(3, 3, 5, mat=J2_STEEL, lz=4.0, load=(60.0, 0.0, -10.0))
    mass = np.full((spec.nn, 3), 0.02)
    gamma, beta, dt, alphaM = 0.5, 0.25, 0.01, 0.8
    (c1, c2, c3), (a1, a2, a3, a4) = newmark_coeffs(gamma, beta, dt)
    O = OracleBackend(spec, 1, 1); O.set_mass(spec.node_tags, mass)
    O.L.orc_set_alphaM.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_double]; O.L.orc_set_alphaM(O.h, alphaM)
    D = xb.DeviceModel.from_spec(spec, 1, 1) if False else None
    D = xb.DeviceModel(3, 3)
    D.add_nodes(spec.node_tags, spec.crd); D.fix(spec.fix[:, 0], spec.fix[:, 1]); D.nd_material(1, *spec.materials[0][1:])
    grp = spec.groups[0]; D.add_elements(grp.kind, grp.tags, grp.conn, grp.mat, grp.par)
    D.add_nodal_loads(spec.loads[:, 0].astype(np.int32), spec.loads[:, 1:]); D.set_mass(spec.node_tags, mass)
    D.set_rayleigh_alpha_m(alphaM); D.setup(1, 1); D.to_device(0)
    ptr, idx = O.csr(); neq = O.neq

    def run(M):
        hist = []; t = 0.0
        for s in range(12):
            t += dt
            M.set_transient(c1, c2, c3); M.newmark_predict(a1, a2, a3, a4)
            M.apply_load(min(t, 0.06) / 0.06)                       # ramp, then hold: free vibration + yielding
            M.incr_response(np.zeros(neq), 1.0, c2, c3)
            B = M.form_unbalance(); norms = []
            for it in range(20):
                A = M.form_tangent()
                dU = spla.spsolve(sp.csr_matrix((A, idx, ptr), shape=(neq, neq)).tocsc(), B)
                M.incr_response(dU, 1.0, c2, c3)
                B = M.form_unbalance()
                norms.append(float(np.linalg.norm(dU)))
                if norms[-1] <= 1e-9:
                    break
            hist.append(norms); M.commit()
        return hist, M.vel_accel()

    ho, (vo, ao) = run(O)
    hd, (vd, ad) = run(D)
    assert [len(h) for h in ho] == [len(h) for h in hd]
    assert max(len(h) for h in ho) >= 3
    assert relerr(vd, vo) < 1e-8 and relerr(ad, ao) < 1e-8


def test_displacement_control_device_vs_golden_reference_history():
    """BASELINE configs[0] (Ex2b cantilever pushover, fibre section) and configs[2] in small (J2 brick
    column): `integrator DisplacementControl` + Newton as run by the reference's own classes (golden)
    against the same algorithm driving the device path: identical iteration counts on every step,
    same load-factor history, same final displacements."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from golden_cases import DISPCONTROL_CASES
    from modelspec import disp_control
    for name, (mk, numberer, soe, node, dof, incr, nsteps, tol, max_iter) in DISPCONTROL_CASES.items():
        g = np.load(os.path.join(GOLD, name + ".npz"))
        spec = mk()
        D = xb.DeviceModel.from_spec(spec, numberer, soe).to_device(0)
        ids = D.ids()
        assert np.array_equal(ids, g["ids"])
        ptr, idx = D.pattern(); neq = D.neq

        def solve(A, b):
            M = sp.csr_matrix((A, idx, ptr), shape=(neq, neq))
            return spla.spsolve((M.T if soe == 0 else M).tocsc(), b)

        ctrl = ids[list(spec.node_tags).index(int(g["node"])), dof]
        hist, lam = disp_control(D, solve, ctrl, incr, nsteps, tol, max_iter, True)
        assert [len(h) for h in hist] == g["iters"].tolist(), name
        assert relerr(lam, g["lam"]) < 1e-8
        assert relerr(D.trial_disp(), g["u"]) < 1e-8
