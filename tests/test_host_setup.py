"""CPU: the product's host side (xara_b200/csrc/host_model.cpp, through the C-ABI) against the
oracle -- DOF numbering, FE order, sparse pattern and scatter maps are BIT-EXACT -- plus the
library/ABI surface.  No device call is made here."""
import os
import re

import numpy as np
import pytest

import xara_b200 as xb
from golden_cases import CASES
from modelspec import (ELASTIC, J2_STEEL, OracleBackend, RefBackend, have_ref, brick_block, brick_periodic_equaldof, frame2d, frame2d_diaphragm_equaldof,
                       quad_plane, soil_column_equaldof)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "xara_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(xb_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert len(declared) >= 35
    import ctypes
    L = ctypes.CDLL(xb.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(xb.EXPORTS), declared ^ set(xb.EXPORTS)


def relabel(spec, seed):
    """random gappy node tags and shuffled element tags: tag->index maps and FE ordering"""
    rng = np.random.default_rng(seed)
    nn = spec.nn
    new = rng.choice(np.arange(1, 7 * nn), nn, replace=False).astype(np.int32)   # new tag of old node i
    order = np.argsort(new)
    spec.node_tags, spec.crd = new[order], spec.crd[order]
    for g in spec.groups:
        g.conn = new[g.conn - 1].astype(np.int32)
        g.tags = rng.permutation(g.tags * 3 + 1).astype(np.int32)
    spec.fix[:, 0] = new[spec.fix[:, 0] - 1]
    spec.loads[:, 0] = new[spec.loads[:, 0].astype(int) - 1]
    return spec


SPECS = {
    "brick": lambda: brick_block(4, 3, 2, distort=0.2),
    "brick_relabel": lambda: relabel(brick_block(3, 3, 3), 5),
    "quad_relabel": lambda: relabel(quad_plane(7, 4, distort=0.2), 6),
    "brick_sliver": lambda: brick_block(9, 1, 1),
    "frame2d": lambda: frame2d(3, 4, 2),
    # `equalDOF` (MP_Constraint with an identity matrix): constrained dofs share the retained dof's equation
    "soilcolumn_equaldof": lambda: soil_column_equaldof(7),
    "brick_equaldof": lambda: brick_periodic_equaldof(3, 2, 3),
    "frame2d_equaldof": lambda: frame2d_diaphragm_equaldof(3, 3, 2),
}


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("numberer", [0, 1])
@pytest.mark.parametrize("soe", [0, 1])
def test_numbering_pattern_scatter_bit_exact(name, numberer, soe):
    spec = SPECS[name]()
    O = OracleBackend(spec, numberer, soe)
    D = xb.DeviceModel.from_spec(spec, numberer, soe)
    nd = {0: 24, 1: 8, 2: 6}[spec.groups[0].kind]
    assert D.neq == O.neq and D.nnz == O.nnz
    assert np.array_equal(D.node_tags(), spec.node_tags)
    assert np.array_equal(D.ids(), O.ids())
    ptr, idx = D.pattern(); po, io = O.csr()
    assert np.array_equal(ptr, po) and np.array_equal(idx, io)
    assert np.array_equal(D.element_tags(), O.fe_ids(nd)[0])
    sm = D.scatter_map(0, D.ne, nd)
    for e in range(D.ne):
        assert np.array_equal(sm[e], O.scatter_map(e, nd))


@pytest.mark.parametrize("name", list(CASES))
def test_host_setup_vs_golden(name):
    mk, numberer, soe, _ = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    D = xb.DeviceModel.from_spec(mk(), numberer, soe)
    assert np.array_equal(D.ids(), g["ids"])
    ptr, idx = D.pattern()
    assert np.array_equal(ptr, g["ptr"]) and np.array_equal(idx, g["idx"])


def test_nodes_may_arrive_unsorted_and_in_batches():
    spec = brick_block(2, 2, 2, distort=0.1)
    D0 = xb.DeviceModel.from_spec(spec, 1, 0)
    perm = np.random.default_rng(0).permutation(spec.nn)
    m = xb.DeviceModel(3, 3)
    h = spec.nn // 2
    m.add_nodes(spec.node_tags[perm[:h]], spec.crd[perm[:h]])
    m.add_nodes(spec.node_tags[perm[h:]], spec.crd[perm[h:]])
    m.fix(spec.fix[:, 0], spec.fix[:, 1])
    m.nd_material(1, *spec.materials[0][1:])
    g = spec.groups[0]
    m.add_elements(g.kind, g.tags[::-1], g.conn[::-1], g.mat[::-1], g.par[::-1])
    m.setup(1, 0)
    assert np.array_equal(m.ids(), D0.ids())
    assert all(np.array_equal(a, b) for a, b in zip(m.pattern(), D0.pattern()))
    assert np.array_equal(m.scatter_map(0, m.ne, 24), D0.scatter_map(0, D0.ne, 24))


def test_error_behaviour():
    m = xb.DeviceModel(3, 3)
    m.add_nodes([1, 2], np.zeros((2, 3)))
    with pytest.raises(xb.XaraB200Error):
        m.nd_material(1, 99, [1.0, 2.0])                      # unknown kind
    m.nd_material(1, xb.MAT_ELASTIC_ISOTROPIC, [100.0, 0.3, 0.0])
    with pytest.raises(xb.XaraB200Error):
        m.nd_material(1, xb.MAT_ELASTIC_ISOTROPIC, [100.0, 0.3, 0.0])   # duplicate tag
    with pytest.raises(xb.XaraB200Error):
        m.add_elements(xb.ELE_STDBRICK, [1], [[1, 2, 3, 4, 5, 6, 7, 8]], [7], np.zeros((1, 3)))  # unknown material
    m.add_elements(xb.ELE_STDBRICK, [1], [[1, 2, 3, 4, 5, 6, 7, 8]], [1], np.zeros((1, 3)))
    with pytest.raises(xb.XaraB200Error):
        m.setup(0, 0)                                          # unknown node tags 3..8
    m2 = xb.DeviceModel(3, 3)
    with pytest.raises(xb.XaraB200Error):
        m2.update()                                            # no device phase yet: never a CPU fallback
    with pytest.raises(xb.XaraB200Error):
        xb.DeviceModel(2, 2).add_elements(xb.ELE_STDBRICK, [1], [[1] * 8], [1], np.zeros((1, 3)))
    # FourNodeQuad's own density (par[3]) replaces the material's in the reference (FourNodeQuad.cpp:395-398): refused,
    # not silently dropped
    q = xb.DeviceModel(2, 2)
    q.add_nodes([1, 2, 3, 4], np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]]))
    q.nd_material(1, xb.MAT_ELASTIC_ISOTROPIC, [100.0, 0.3, 0.0])
    with pytest.raises(xb.XaraB200Error, match="density"):
        q.add_elements(xb.ELE_FOURNODEQUAD, [1], [[1, 2, 3, 4]], [1], np.array([[1.0, 0, 0, 2.5, 0, 0]]))   # thick type pressure rho b1 b2
    q.add_elements(xb.ELE_FOURNODEQUAD, [1], [[1, 2, 3, 4]], [1], np.array([[1.0, 0, 0, 0.0, 0, 0]]))
    assert q.set_option("ranged_tangent", 0) is q
    with pytest.raises(xb.XaraB200Error):
        q.set_option("no_such_option", 1)


def test_beam_edges_refused_on_the_host():
    """what the device path does not cover is refused when the model is described, not computed wrongly: a corotational
    transformation in 3D, a partial uniform load with aOverL >= bOverL or a second one on the same element, stiffness-
    proportional Rayleigh terms on corotational beams"""
    from modelspec import frame2d, frame3d, with_corot, with_beam_partial_loads
    sp3 = frame3d(1, 1, 1)
    for g in sp3.groups: g.par[:, 6] = 2.0
    with pytest.raises(xb.XaraB200Error, match="[Cc]orotational"):
        xb.DeviceModel.from_spec(sp3, setup=False)
    sp = with_beam_partial_loads(with_corot(frame2d(1, 1, 1)), seed=1)
    m = xb.DeviceModel.from_spec(sp, setup=False)
    t = [sp.beam_partial_loads[0][0]]
    with pytest.raises(xb.XaraB200Error, match="one partial"):
        m.add_beam_partial_loads(t, np.array([[-0.1, -0.1, 0, 0, 0.2, 0.8, 0, 0]]))
    col = [int(x) for x in sp.groups[0].tags if int(x) not in {q[0] for q in sp.beam_partial_loads}][:1]
    with pytest.raises(xb.XaraB200Error, match="aOverL"):
        m.add_beam_partial_loads(col, np.array([[-0.1, -0.1, 0, 0, 0.8, 0.2, 0, 0]]))
    with pytest.raises(xb.XaraB200Error, match="corotational"):
        m.set_rayleigh(0.1, 0.002, 0.0, 0.0)
    m.set_rayleigh(0.1, 0.0, 0.0, 0.0)          # mass-proportional damping alone is fine


@pytest.mark.skipif(xb.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    D = xb.DeviceModel.from_spec(brick_block(1, 1, 1), 0, 0)
    with pytest.raises(xb.XaraB200Error):
        D.to_device(0)
    with pytest.raises(xb.XaraB200Error):
        D.form_tangent()


def test_equal_dof_edge_cases():
    # a retained node with no element of its own, a constrained node tied onto a fixed dof, and refusals
    spec = brick_block(2, 1, 1)
    spec.node_tags = np.append(spec.node_tags, 100).astype(np.int32)
    spec.crd = np.vstack([spec.crd, [5.0, 5.0, 5.0]])
    spec.equal_dofs = [(100, int(spec.node_tags[-2]), [0, 2]), (1, int(spec.node_tags[-3]), [1])]   # node 1 is fixed
    for numberer in (0, 1):
        for soe in (0, 1):
            O = OracleBackend(spec, numberer, soe); D = xb.DeviceModel.from_spec(spec, numberer, soe)
            ids = D.ids()
            assert np.array_equal(ids, O.ids())
            assert np.array_equal(ids[-1, [0, 2]], ids[-2, [0, 2]]) and ids[-3, 1] == -1
            assert all(np.array_equal(a, b) for a, b in zip(D.pattern(), O.csr()))
            sm = D.scatter_map(0, D.ne, 24)
            for e in range(D.ne):
                assert np.array_equal(sm[e], O.scatter_map(e, 24))
    chain = brick_block(2, 1, 1)
    chain.equal_dofs = [(7, 8, [0]), (8, 9, [0])]            # node 8's dof is retained AND constrained
    with pytest.raises(xb.XaraB200Error):
        xb.DeviceModel.from_spec(chain, 0, 0)
    from modelspec import quad_plane_stress_pressure
    mixed = quad_plane_stress_pressure(3, 3, 1, 1.0, mat=J2_STEEL)
    mixed.groups[0].par[::2, 1] = 0
    with pytest.raises(xb.XaraB200Error):      # J2 quads: one plane type per batch (J2PlaneStress keeps a state of its own)
        xb.DeviceModel.from_spec(mixed, 0, 0)
    m = xb.DeviceModel(3, 3)
    m.add_nodes([1, 2], np.zeros((2, 3)))
    with pytest.raises(xb.XaraB200Error):
        m.equal_dof(1, 1, [0])
    with pytest.raises(xb.XaraB200Error):
        m.equal_dof(1, 2, [3])


def ragged_spec(seed):
    """a random subset of a brick block's elements, random extra fixes and (two seeds in three) `equalDOF` ties"""
    rng = np.random.default_rng(100 + seed)
    spec = brick_block(4, 3, 3, distort=0.1, seed=seed)
    g = spec.groups[0]
    keep = rng.random(spec.ne) < rng.uniform(0.35, 0.9)
    keep[rng.integers(spec.ne)] = True
    order = rng.permutation(np.where(keep)[0])
    g.tags, g.conn, g.mat, g.par = (g.tags[order] * 5 + 2).astype(np.int32), g.conn[order], g.mat[order], g.par[order]
    have = {(int(t), int(d)) for t, d in spec.fix}
    extra = [(int(t), int(rng.integers(3))) for t in rng.choice(spec.node_tags, 4, replace=False)]
    extra = [td for td in extra if td not in have]
    if extra:
        spec.fix = np.vstack([spec.fix, extra]).astype(np.int32)
    if seed % 3:        # ties: distinct retained / constrained nodes, no chains
        nodes = rng.choice(spec.node_tags, 8, replace=False)
        spec.equal_dofs = [(int(nodes[2 * i]), int(nodes[2 * i + 1]), sorted(rng.choice(3, rng.integers(1, 4), replace=False).tolist()))
                           for i in range(4)]
    return spec


@pytest.mark.parametrize("seed", range(12))
def test_ragged_random_meshes_bit_exact(seed):
    """ragged input: a random subset of a brick block's elements (holes, nodes left without any element, disconnected
    pieces), random extra fixes, random `equalDOF` ties (free-to-free, onto fixed dofs, onto element-less nodes), shuffled
    element order and gappy tags -- numbering, pattern and scatter maps stay bit-exact against the oracle"""
    spec = ragged_spec(seed)
    for numberer in (0, 1):
        for soe in (0, 1):
            O = OracleBackend(spec, numberer, soe)
            D = xb.DeviceModel.from_spec(spec, numberer, soe)
            assert D.neq == O.neq and D.nnz == O.nnz
            assert np.array_equal(D.ids(), O.ids())
            assert all(np.array_equal(a, b) for a, b in zip(D.pattern(), O.csr()))
            assert np.array_equal(D.element_tags(), O.fe_ids(24)[0])
            sm = D.scatter_map(0, D.ne, 24)
            for e in range(D.ne):
                assert np.array_equal(sm[e], O.scatter_map(e, 24))
            fixed = {(int(t), int(d)) for t, d in spec.fix}
            quiet = not any((c, d) in fixed for _, c, dofs in spec.equal_dofs for d in dofs)   # (the reference warns on those)
            if have_ref() and soe == 1 and quiet:   # ... and the oracle against the reference's own handler / numberer / SOE
                R = RefBackend(spec, numberer, soe)
                assert np.array_equal(O.ids(), R.ids())
                assert all(np.array_equal(a, b) for a, b in zip(O.csr(), R.csr()))


def test_all_fixed_and_isolated_nodes():
    spec = brick_block(1, 1, 1)
    spec.fix = np.array([(t, d) for t in spec.node_tags for d in range(3)], np.int32)
    D = xb.DeviceModel.from_spec(spec, 0, 0)
    assert D.neq == 0 and D.nnz == 0
    spec = brick_block(1, 1, 1)
    spec.node_tags = np.append(spec.node_tags, 100).astype(np.int32)
    spec.crd = np.vstack([spec.crd, [5.0, 5.0, 5.0]])
    for numberer in (0, 1):
        O = OracleBackend(spec, numberer, 1); D = xb.DeviceModel.from_spec(spec, numberer, 1)
        assert np.array_equal(D.ids(), O.ids())
        assert all(np.array_equal(a, b) for a, b in zip(D.pattern(), O.csr()))


@pytest.mark.parametrize("soe", [2, 3, 4])
@pytest.mark.parametrize("numberer", [0, 1])
def test_band_profile_umfpack_layout_and_scatter_maps_bit_exact(numberer, soe):
    """`system BandGeneral | ProfileSPD | Umfpack`: layout (numSubD / numSuperD, iDiagLoc, Ap / Ai) and the addA location of
    every element-matrix entry against the oracle's restatement of BandGenLinSOE / ProfileSPDLinSOE (pinned to the live
    classes in tests/test_oracle.py); Umfpack's Ap / Ai are SparseGenCol's colStartA / rowA"""
    for spec in (brick_block(4, 3, 2, distort=0.2, seed=5), quad_plane(6, 4, distort=0.2, seed=6), frame2d(2, 3, 2),
                 soil_column_equaldof(6), brick_periodic_equaldof(3, 2, 2)):
        nd = {0: 24, 1: 8, 2: 6, 3: 12}[spec.groups[0].kind]
        O = OracleBackend(spec, numberer, soe)
        D = xb.DeviceModel.from_spec(spec, numberer, soe)
        assert D.neq == O.neq and D.a_size == O.a_size and np.array_equal(D.ids(), O.ids())
        assert all(np.array_equal(a, b) for a, b in zip(D.pattern(), O.csr()))
        if soe == 2:
            assert D.band() == O.band()
        if soe == 3:
            assert np.array_equal(D.profile(), O.profile())
        for e in range(O.ne):
            assert np.array_equal(D.scatter_map(e, e + 1, nd).reshape(nd, nd), O.scatter_map(e, nd))


@pytest.mark.parametrize("numberer,soe", [(0, 0), (1, 1), (1, 2)])
def test_mixed_ndf_soil_frame_bit_exact(numberer, soe):
    """a FourNodeQuad soil layer on 2-dof nodes carrying a forceBeamColumn frame on 3-dof nodes (`equalDOF` at the column
    bases): DOF ids, pattern and every element's addA locations against the oracle (pinned to the live reference in
    tests/test_oracle.py::test_mixed_ndf_soil_frame_vs_live_reference); elements on nodes of the wrong size are refused"""
    from modelspec import soil_frame_2d
    spec = soil_frame_2d()
    O = OracleBackend(spec, numberer, soe)
    D = xb.DeviceModel.from_spec(spec, numberer, soe)
    assert D.neq == O.neq and np.array_equal(D.ids(), O.ids())
    assert all(np.array_equal(a, b) for a, b in zip(D.pattern(), O.csr()))
    nq = len(spec.groups[0].tags)
    for e in range(O.ne):
        nd = 8 if e < nq else 6
        assert np.array_equal(D.scatter_map(e, e + 1, nd).reshape(nd, nd), O.scatter_map(e, nd))
    bad = soil_frame_2d(); bad.node_ndf = {}                      # quads on 3-dof nodes: FourNodeQuad.cpp:133-139 says no
    with pytest.raises(xb.XaraB200Error, match="number of dofs"):
        xb.DeviceModel.from_spec(bad, numberer, soe)
