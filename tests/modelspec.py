"""Model descriptions shared by every backend the tests compare.

A ``ModelSpec`` is the flat, array-of-arrays form of what the reference's model
commands (``node``, ``fix``, ``nDMaterial``, ``element stdBrick|quad``, ``load``)
build in a ``Domain``.  Three backends consume it:

* ``RefBackend``    -- the reference's own classes (oracle/_ref/libref_harness.so)
* ``OracleBackend`` -- the C restatement (oracle/liboracle.so)
* ``xara_b200.DeviceModel`` -- the product (CUDA, through the C-ABI)

Only tests/, bench.py's CPU legs and __graft_entry__.smoke() import this.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
METIS_SO = os.path.join(ROOT, "oracle", "_ref", "libmetis_ref.so")
GLUE_SO = os.path.join(ROOT, "oracle", "_ref", "libref_glue.so")

MAT_ELASTIC, MAT_J2 = 0, 1
ELE_BRICK, ELE_QUAD, ELE_FBC2D, ELE_FBC3D = 0, 1, 2, 3
UNI_STEEL02, UNI_CONCRETE02, UNI_STEEL01, UNI_ELASTIC, UNI_CONCRETE01, UNI_ELASTICPP = 0, 1, 2, 3, 4, 5
SEC_P, SEC_MZ = 2, 1     # SectionForceDeformation.h response codes
ND_3D, ND_PLANE_STRAIN, ND_PLANE_STRESS = 0, 1, 2
NUMBERER_PLAIN, NUMBERER_RCM = 0, 1
SOE_CSC, SOE_CSR = 0, 1


@dataclass
class ElementGroup:
    kind: int                 # ELE_BRICK | ELE_QUAD
    tags: np.ndarray          # [ne] int32
    conn: np.ndarray          # [ne, nen] int32 node TAGS
    mat: np.ndarray           # [ne] int32 material tags
    par: np.ndarray           # [ne, 8] float64 (brick: b1,b2,b3; quad: thick,type,pressure,rho,b1,b2)


@dataclass
class ModelSpec:
    ndm: int
    ndf: int
    node_tags: np.ndarray     # [nn] int32 ascending
    crd: np.ndarray           # [nn, ndm]
    fix: np.ndarray           # [nfix, 2] (node tag, dof)
    materials: list           # (tag, kind, params[<=8])
    groups: list = field(default_factory=list)
    loads: np.ndarray = None  # [nload, 1+ndf] (node tag, values...)
    uniaxials: list = field(default_factory=list)   # (tag, kind, params)
    equal_dofs: list = field(default_factory=list)  # `equalDOF`: (retained node tag, constrained node tag, [dofs], 0-based)
    sections: list = field(default_factory=list)    # (tag, y[nf], A[nf], uniaxial tags[nf])
    beam_loads: list = field(default_factory=list)  # `eleLoad -beamUniform`: (element tag, wy, wz, wa) in the Linear pattern
    beam_point_loads: list = field(default_factory=list)  # `eleLoad -beamPoint`: (element tag, Py, Pz, N, xL)
    beam_partial_loads: list = field(default_factory=list)  # partial `-beamUniform`: (element tag, wya, wyb, waa, wab, aOverL, bOverL, wza, wzb)
    beam_integration: int = 0          # forceBeamColumn -integration: 0 Lobatto, 1 Legendre, 2 Radau, 3 NewtonCotes, 4 Trapezoidal
    beam_rules: tuple = None           # (element tags, xi [n][nip], wt [n][nip]) as the reference's BeamIntegration returns them
    node_ndf: dict = field(default_factory=dict)    # node tag -> dofs, for nodes created under another `model -ndf` (< the model's ndf)

    @property
    def nn(self):
        return len(self.node_tags)

    @property
    def ne(self):
        return sum(len(g.tags) for g in self.groups)


J2_STEEL = (MAT_J2, [166.67e3, 76.92e3, 250.0, 400.0, 16.93, 500.0, 0.0])   # K G sig0 sigInf delta H eta
ELASTIC = (MAT_ELASTIC, [1000.0, 0.25, 6.75])                                # E nu rho (Verification/Plane/PlaneStrain.tcl)


def brick_block(nx, ny, nz, mat=J2_STEEL, lx=1.0, ly=1.0, lz=1.0, distort=0.0, seed=0,
                body=(0.0, 0.0, 0.0), fix_face="z0", load=None):
    """nx*ny*nz stdBrick block, node tags 1.. in x-fastest order, base fixed."""
    rng = np.random.default_rng(seed)
    X, Y, Z = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    idx = (X + (nx + 1) * (Y + (ny + 1) * Z))
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    crd = np.zeros((nn, 3))
    crd[idx.ravel(), 0] = X.ravel() * lx / nx
    crd[idx.ravel(), 1] = Y.ravel() * ly / ny
    crd[idx.ravel(), 2] = Z.ravel() * lz / nz
    if distort:
        h = min(lx / nx, ly / ny, lz / nz)
        crd += distort * h * (rng.random(crd.shape) - 0.5)
    ex, ey, ez = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ex, ey, ez = [a.transpose(2, 1, 0).ravel() for a in (ex, ey, ez)]

    def nid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k) + 1

    conn = np.stack([nid(ex, ey, ez), nid(ex + 1, ey, ez), nid(ex + 1, ey + 1, ez), nid(ex, ey + 1, ez),
                     nid(ex, ey, ez + 1), nid(ex + 1, ey, ez + 1), nid(ex + 1, ey + 1, ez + 1),
                     nid(ex, ey + 1, ez + 1)], axis=1).astype(np.int32)
    ne = len(conn)
    par = np.zeros((ne, 8)); par[:, :3] = body
    base = np.where(idx[:, :, 0].ravel() >= 0)[0]
    base_tags = (idx[:, :, 0].ravel() + 1)
    fix = np.array([(t, d) for t in base_tags for d in range(3)], dtype=np.int32).reshape(-1, 2)
    top_tags = idx[:, :, nz].ravel() + 1
    if load is None:
        load = (0.3, 0.0, -1.0)
    loads = np.array([[t, *load] for t in top_tags], dtype=np.float64)
    return ModelSpec(3, 3, np.arange(1, nn + 1, dtype=np.int32), crd, fix, [(1, *mat)],
                     [ElementGroup(ELE_BRICK, np.arange(1, ne + 1, dtype=np.int32), conn,
                                   np.ones(ne, np.int32), par)], loads)


def quad_plane(nx, ny, mat=ELASTIC, lx=40.0, ly=10.0, thick=1.0, distort=0.0, seed=0, body=(0.0, 0.0)):
    """nx*ny FourNodeQuad plane-strain mesh (Verification/Plane/PlaneStrain.tcl shape), left edge fixed."""
    rng = np.random.default_rng(seed)
    X, Y = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    idx = X + (nx + 1) * Y
    nn = (nx + 1) * (ny + 1)
    crd = np.zeros((nn, 2))
    crd[idx.ravel(), 0] = X.ravel() * lx / nx
    crd[idx.ravel(), 1] = Y.ravel() * ly / ny
    if distort:
        crd += distort * min(lx / nx, ly / ny) * (rng.random(crd.shape) - 0.5)
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ex, ey = ex.T.ravel(), ey.T.ravel()

    def nid(i, j):
        return i + (nx + 1) * j + 1

    conn = np.stack([nid(ex, ey), nid(ex + 1, ey), nid(ex + 1, ey + 1), nid(ex, ey + 1)], axis=1).astype(np.int32)
    ne = len(conn)
    par = np.zeros((ne, 8)); par[:, 0] = thick; par[:, 1] = 0; par[:, 4:6] = body
    left = idx[0, :].ravel() + 1
    fix = np.array([(t, d) for t in left for d in range(2)], dtype=np.int32).reshape(-1, 2)
    right = idx[nx, :].ravel() + 1
    loads = np.array([[t, 0.5, -1.0] for t in right], dtype=np.float64)
    return ModelSpec(2, 2, np.arange(1, nn + 1, dtype=np.int32), crd, fix, [(1, *mat)],
                     [ElementGroup(ELE_QUAD, np.arange(1, ne + 1, dtype=np.int32), conn,
                                   np.ones(ne, np.int32), par)], loads)


STEEL02 = (UNI_STEEL02, [60.0, 29000.0, 0.01, 18.0, 0.925, 0.15, 0.0, 1.0, 0.0, 1.0, 0.0])   # Fy E0 b R0 cR1 cR2 a1..a4 sigInit
CONCRETE02_CORE = (UNI_CONCRETE02, [-6.0, -0.004, -5.0, -0.014, 0.1, 0.6, 300.0])            # fc epsc0 fcu epscu rat ft Ets
CONCRETE02_COVER = (UNI_CONCRETE02, [-5.0, -0.002, 0.0, -0.006, 0.1, 0.5, 250.0])


def rc_section(tag=1, h=24.0, b=15.0, cover=1.5, nf_core=10, nf_cover=2, As=0.6):
    """the classic RC fibre section (OpenSees example 'RCFrameGravity'): confined core, unconfined cover,
    three layers of steel -- as (y, A, uniaxial tag) fibres; uniaxial tags 1 core, 2 cover, 3 steel"""
    y, A, m = [], [], []
    y1, z1 = h / 2.0, b / 2.0
    hc = h - 2 * cover
    for i in range(nf_core):                      # core patch
        y.append(-hc / 2 + (i + 0.5) * hc / nf_core); A.append((b - 2 * cover) * hc / nf_core); m.append(1)
    for i in range(nf_core):                      # side cover
        y.append(-hc / 2 + (i + 0.5) * hc / nf_core); A.append(2 * cover * hc / nf_core); m.append(2)
    for sgn in (-1, 1):                           # top / bottom cover
        for i in range(nf_cover):
            y.append(sgn * (hc / 2 + (i + 0.5) * cover / nf_cover)); A.append(b * cover / nf_cover); m.append(2)
    for yy, n in ((y1 - cover, 3), (0.0, 2), (-(y1 - cover), 3)):   # steel layers
        for _ in range(n):
            y.append(yy); A.append(As); m.append(3)
    return (tag, np.array(y), np.array(A), np.array(m, np.int32))


def frame2d(nbay=2, nstory=3, ndiv=2, nip=5, bay=360.0, story=144.0, max_iters=10, tol=1e-12, lateral=10.0, gravity=-60.0):
    """2D RC moment frame of forceBeamColumn elements with fibre sections (Steel02 + Concrete02),
    every member split in ndiv elements; bases fixed; gravity on the floor nodes + lateral load at the roof"""
    pts = {}
    def node(x, y):
        key = (round(x, 6), round(y, 6))
        if key not in pts:
            pts[key] = len(pts) + 1
        return pts[key]
    conn = []
    for i in range(nbay + 1):
        for j in range(nstory):
            for d in range(ndiv):
                conn.append((node(i * bay, j * story + d * story / ndiv), node(i * bay, j * story + (d + 1) * story / ndiv)))
    for j in range(1, nstory + 1):
        for i in range(nbay):
            for d in range(ndiv):
                conn.append((node(i * bay + d * bay / ndiv, j * story), node(i * bay + (d + 1) * bay / ndiv, j * story)))
    nn = len(pts)
    crd = np.zeros((nn, 2))
    for (x, y), t in pts.items():
        crd[t - 1] = (x, y)
    conn = np.array(conn, np.int32)
    ne = len(conn)
    par = np.zeros((ne, 8)); par[:, 0] = nip; par[:, 1] = max_iters; par[:, 2] = tol
    fix = np.array([(t, d) for (x, y), t in pts.items() if y == 0.0 for d in range(3)], np.int32).reshape(-1, 2)
    loads = []
    for (x, y), t in pts.items():
        if y > 0 and abs(y / story - round(y / story)) < 1e-9 and abs(x / bay - round(x / bay)) < 1e-9:
            loads.append([t, lateral if (x == 0.0 and round(y / story) == nstory) else 0.0, gravity, 0.0])
    return ModelSpec(2, 3, np.arange(1, nn + 1, dtype=np.int32), crd, fix, [],
                     [ElementGroup(ELE_FBC2D, np.arange(1, ne + 1, dtype=np.int32), conn, np.ones(ne, np.int32), par)],
                     np.array(loads), uniaxials=[(1, *CONCRETE02_CORE), (2, *CONCRETE02_COVER), (3, *STEEL02)],
                     sections=[rc_section(1)])


def rc_section3d(tag=1, h=24.0, b=18.0, cover=1.5, ny=6, nz=4, As=0.6, GJ=2.0e6):
    """3D RC fibre section (`section Fiber tag -GJ gj`): confined core patch ny x nz, four cover strips,
    eight bars -- as (tag, y, A, uniaxial tag, z, GJ); uniaxial tags 1 core, 2 cover, 3 steel"""
    y, z, A, m = [], [], [], []
    hc, bc = h - 2 * cover, b - 2 * cover
    for i in range(ny):
        for j in range(nz):
            y.append(-hc / 2 + (i + 0.5) * hc / ny); z.append(-bc / 2 + (j + 0.5) * bc / nz)
            A.append(hc * bc / (ny * nz)); m.append(1)
    for sgn in (-1, 1):
        for j in range(nz):       # top / bottom cover strips (full width split in nz)
            y.append(sgn * (h - cover) / 2); z.append(-b / 2 + (j + 0.5) * b / nz); A.append(cover * b / nz); m.append(2)
        for i in range(ny):       # side cover strips
            y.append(-hc / 2 + (i + 0.5) * hc / ny); z.append(sgn * (b - cover) / 2); A.append(cover * hc / ny); m.append(2)
    for yy in (-hc / 2, 0.0, hc / 2):
        for zz in (-bc / 2, 0.0, bc / 2):
            if yy == 0.0 and zz == 0.0:
                continue
            y.append(yy); z.append(zz); A.append(As); m.append(3)
    return (tag, np.array(y), np.array(A), np.array(m, np.int32), np.array(z), GJ)


def soil_frame_2d(nbay=2, nstory=2, ndiv=1, per_bay=3, ny=4, depth=240.0, mat=J2_STEEL, distort=0.1, seed=7, lateral=8.0, gravity=-30.0):
    """BASELINE configs[4] in 2D, as a real mixed-ndf Domain: a FourNodeQuad soil layer whose nodes were created under
    `model -ndf 2` carrying an RC frame of forceBeamColumn elements on 3-dof nodes (`model -ndf 3`); every column base
    sits on a soil surface node and follows it in both translations (`equalDOF soil base 1 2`), its rotation fixed.
    Soil base fixed, soil sides free; loads on the frame's floor nodes."""
    fr = frame2d(nbay, nstory, ndiv, lateral=lateral, gravity=gravity)
    bay = 360.0
    nx = nbay * per_bay + 2 * per_bay                      # one bay of free field on either side
    lx = (nbay + 2) * bay
    so = quad_plane(nx, ny, mat=mat, lx=lx, ly=depth, distort=0.0, body=(0.0, -0.002))
    rng = np.random.default_rng(seed)
    scrd = so.crd.copy()
    inner = (scrd[:, 1] > 1e-9) & (scrd[:, 1] < depth - 1e-9)
    scrd[inner] += distort * min(lx / nx, depth / ny) * (rng.random((inner.sum(), 2)) - 0.5)
    ns = so.nn
    off_n, off_e = ns, so.ne                               # frame node / element tags follow the soil's
    fcrd = fr.crd + np.array([bay, depth])                 # the frame stands on the soil surface, one bay in
    tags = np.concatenate([so.node_tags, fr.node_tags + off_n]).astype(np.int32)
    crd = np.vstack([scrd, fcrd])
    fix = [(int(t), d) for t in so.node_tags[scrd[:, 1] < 1e-9] for d in range(2)]
    eq = []
    for t, (x, y) in zip(fr.node_tags, fr.crd):
        if y == 0.0:                                       # a column base: tie to the soil node under it, fix the rotation
            k = int(np.argmin(np.abs(scrd[:, 0] - (x + bay)) + np.abs(scrd[:, 1] - depth)))
            assert abs(scrd[k, 0] - (x + bay)) < 1e-6 and abs(scrd[k, 1] - depth) < 1e-6
            eq.append((int(so.node_tags[k]), int(t) + off_n, [0, 1]))
            fix.append((int(t) + off_n, 2))
    g0, g1 = so.groups[0], fr.groups[0]
    groups = [ElementGroup(ELE_QUAD, g0.tags, g0.conn, g0.mat, g0.par),
              ElementGroup(ELE_FBC2D, g1.tags + off_e, g1.conn + off_n, g1.mat, g1.par)]
    loads = fr.loads.copy(); loads[:, 0] += off_n
    spec = ModelSpec(2, 3, tags, crd, np.array(fix, np.int32).reshape(-1, 2), so.materials, groups, loads,
                     uniaxials=fr.uniaxials, equal_dofs=eq, sections=fr.sections)
    spec.node_ndf = {int(t): 2 for t in so.node_tags}
    return spec


def steel01_elastic_frame(dim, bars="steel01"):
    """an RC frame whose bars are Steel01 (with isotropic hardening), whose core is Concrete01 and whose cover is a bilinear
    Elastic material: the uniaxial kinds of BASELINE configs[0], and the concrete most RC examples use, as fibres of
    FiberSection2d / FiberSection3d"""
    spec = frame2d(2, 2, 2, lateral=20.0) if dim == 2 else frame3d(1, 1, 2, ndiv=2, lateral=(18.0, 10.0))
    uni = dict((t, (k, p)) for t, k, p in spec.uniaxials)
    uni[1] = (UNI_CONCRETE01, (-6.0, -0.004, -5.0, -0.014))             # core: Kent-Scott-Park, no tension
    uni[2] = (UNI_ELASTIC, (2500.0, 0.0, 3600.0))                       # cover: softer in tension
    uni[3] = (UNI_STEEL01, (60.0, 29000.0, 0.015, 0.02, 30.0, 0.02, 30.0))
    if bars == "elasticpp":       # `uniaxialMaterial ElasticPP E epsyP epsyN eps0`: bars without hardening, weaker in compression, pre-strained
        uni[3] = (UNI_ELASTICPP, (29000.0, 60.0 / 29000.0, -50.0 / 29000.0, 1.0e-4))
    spec.uniaxials = [(t, *uni[t]) for t in sorted(uni)]
    return spec


def with_beam_integration(spec, kind):
    """`-integration Legendre | Radau | NewtonCotes | Trapezoidal` (kind 1..4) on every forceBeamColumn: the reference
    backend builds the elements with that BeamIntegration class; oracle and device get the section locations and weights
    the class returns (ref_beam_rule), which is what the binding reads out of the element"""
    L = ctypes.CDLL(REF_SO)
    L.ref_beam_rule.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    tags, xis, wts = [], [], []
    for g in spec.groups:
        if g.kind not in (ELE_FBC2D, ELE_FBC3D): continue
        nip = int(g.par[0, 0])
        for t, c in zip(g.tags, g.conn):
            idx = [list(spec.node_tags).index(int(q)) for q in c]
            Lel = float(np.linalg.norm(spec.crd[idx[1]] - spec.crd[idx[0]]))
            xi = np.zeros(nip); wt = np.zeros(nip)
            assert L.ref_beam_rule(kind, nip, Lel, _p(xi), _p(wt)) == 0
            tags.append(int(t)); xis.append(xi); wts.append(wt)
    spec.beam_integration = kind
    spec.beam_rules = (np.array(tags, np.int32), np.array(xis), np.array(wts))
    return spec


def with_beam_point_loads(spec, P=-6.0, seed=0):
    """`eleLoad -beamPoint Py [Pz] xL N` on every other forceBeamColumn (xL drawn in [0.15, 0.85], so that some points
    fall between, some beyond, the Lobatto sections), a small axial component along"""
    rng = np.random.default_rng(seed)
    out = []
    for g in spec.groups:
        if g.kind not in (ELE_FBC2D, ELE_FBC3D): continue
        for t in g.tags[::2]:
            out.append((int(t), P * rng.uniform(0.5, 1.5), (0.4 * P * rng.uniform(0.5, 1.5)) if g.kind == ELE_FBC3D else 0.0,
                        0.1 * P * rng.uniform(-1, 1), float(rng.uniform(0.15, 0.85))))
    spec.beam_point_loads = out
    return spec


def with_beam_rho(spec, rho=2.0e-3):
    """`element forceBeamColumn ... -mass rho` (mass per unit length; element parameter 4 in 2D, 7 in 3D)"""
    for g in spec.groups:
        if g.kind == ELE_FBC2D: g.par[:, 4] = rho
        elif g.kind == ELE_FBC3D: g.par[:, 7] = rho
    return spec


def with_beam_partial_loads(spec, w=-0.12, seed=0):
    """a trapezoidal `eleLoad -beamUniform ... aOverL bOverL` (Beam2d / Beam3dPartialUniformLoad) over a random part of every
    girder -- section points fall before, inside and behind the loaded stretch.  Entries: (element tag, wy_a, wy_b, wAxial_a,
    wAxial_b, aOverL, bOverL, wz_a, wz_b)"""
    rng = np.random.default_rng(seed)
    ix = {int(t): i for i, t in enumerate(spec.node_tags)}
    up = spec.ndm - 1
    out = []
    for g in spec.groups:
        for t, c in zip(g.tags, g.conn):
            a, b = spec.crd[ix[int(c[0])]], spec.crd[ix[int(c[1])]]
            if abs(a[up] - b[up]) > 1e-9: continue          # a column
            aL = rng.uniform(0.05, 0.45); bL = rng.uniform(0.55, 0.95)
            wa = w * rng.uniform(0.6, 1.4); wb = w * rng.uniform(0.6, 1.4)
            wz = (0.15 * wa, -0.1 * wb) if spec.ndm == 3 else (0.0, 0.0)
            out.append((int(t), wa, wb, 0.05 * wa, -0.03 * wb, aL, bL) + wz)
    spec.beam_partial_loads = out
    return spec


def with_joint_offsets(spec, frac=0.06, seed=0):
    """`geomTransf ... -jntOffset`: rigid end zones on every forceBeamColumn -- each end is moved along the member by a
    random fraction (up to `frac`) of its length, plus a small lateral eccentricity (element parameters 5..8 in 2D:
    dXi dYi dXj dYj)"""
    rng = np.random.default_rng(seed)
    for g in spec.groups:
        if g.kind not in (ELE_FBC2D, ELE_FBC3D): continue
        par = np.zeros((len(g.tags), 14)); par[:, :g.par.shape[1]] = g.par
        for i, c in enumerate(g.conn):
            idx = [list(spec.node_tags).index(int(q)) for q in c]
            d = spec.crd[idx[1]] - spec.crd[idx[0]]
            if g.kind == ELE_FBC2D:
                nrm = np.array([-d[1], d[0]])
                par[i, 5:7] = rng.uniform(0.3, 1.0) * frac * d + rng.uniform(-0.2, 0.2) * frac * nrm
                par[i, 7:9] = -rng.uniform(0.3, 1.0) * frac * d + rng.uniform(-0.2, 0.2) * frac * nrm
            else:        # 3D (parameters 8..13: dXi dYi dZi dXj dYj dZj): along the member plus a small eccentricity
                L = np.linalg.norm(d)
                par[i, 8:11] = rng.uniform(0.3, 1.0) * frac * d + rng.uniform(-0.2, 0.2, 3) * frac * L
                par[i, 11:14] = -rng.uniform(0.3, 1.0) * frac * d + rng.uniform(-0.2, 0.2, 3) * frac * L
        g.par = par
    return spec


def with_pdelta(spec):
    """`geomTransf PDelta` instead of Linear on every forceBeamColumn of the spec (element parameter 3 in 2D, 6 in 3D)"""
    for g in spec.groups:
        if g.kind == ELE_FBC2D: g.par[:, 3] = 1.0
        elif g.kind == ELE_FBC3D: g.par[:, 6] = 1.0
    return spec


def with_corot(spec):
    """`geomTransf Corotational` on every 2D forceBeamColumn of the spec (element parameter 3 = 2)"""
    for g in spec.groups:
        if g.kind == ELE_FBC2D: g.par[:, 3] = 2.0
    return spec


def with_beam_gravity(spec, w=-0.25, axial=0.02, seed=0):
    """`eleLoad -beamUniform` on the horizontal members (girders): transverse w (+-20 % per element), a little axial load,
    and -- 3D -- a small lateral component; columns stay unloaded"""
    rng = np.random.default_rng(seed)
    ix = {int(t): i for i, t in enumerate(spec.node_tags)}
    up = spec.ndm - 1
    loads = []
    for g in spec.groups:
        for t, c in zip(g.tags, g.conn):
            a, b = spec.crd[ix[int(c[0])]], spec.crd[ix[int(c[1])]]
            if abs(a[up] - b[up]) > 1e-9:
                continue                                  # a column
            f = 1.0 + 0.2 * (rng.random() - 0.5)
            loads.append((int(t), w * f, 0.1 * w * f if spec.ndm == 3 else 0.0, axial * f))
    spec.beam_loads = loads
    return spec


def frame3d(nx=1, ny=1, nstory=2, ndiv=1, nip=4, bay=240.0, story=144.0, max_iters=10, tol=1e-12,
            lateral=(8.0, 5.0), gravity=-40.0):
    """3D RC space frame of forceBeamColumn elements (ForceBeamColumn3d, FiberSection3d: Steel02 + Concrete02):
    columns on an (nx+1) x (ny+1) grid, beams in both directions at every floor; bases fixed; gravity on the
    floor nodes, lateral loads (x, y) at the roof corner.  ndm 3, ndf 6."""
    pts = {}
    def node(x, y, z):
        key = (round(x, 6), round(y, 6), round(z, 6))
        if key not in pts:
            pts[key] = len(pts) + 1
        return pts[key]
    conn, vec = [], []
    for i in range(nx + 1):
        for j in range(ny + 1):
            for k in range(nstory):
                for d in range(ndiv):
                    conn.append((node(i * bay, j * bay, k * story + d * story / ndiv), node(i * bay, j * bay, k * story + (d + 1) * story / ndiv)))
                    vec.append((1.0, 0.0, 0.0))          # columns: local x up, vecxz = global X
    for k in range(1, nstory + 1):
        for j in range(ny + 1):
            for i in range(nx):
                for d in range(ndiv):
                    conn.append((node(i * bay + d * bay / ndiv, j * bay, k * story), node(i * bay + (d + 1) * bay / ndiv, j * bay, k * story)))
                    vec.append((0.0, 0.0, 1.0))          # beams: vecxz = global Z
        for i in range(nx + 1):
            for j in range(ny):
                for d in range(ndiv):
                    conn.append((node(i * bay, j * bay + d * bay / ndiv, k * story), node(i * bay, j * bay + (d + 1) * bay / ndiv, k * story)))
                    vec.append((0.0, 0.0, 1.0))
    nn = len(pts)
    crd = np.zeros((nn, 3))
    for key, t in pts.items():
        crd[t - 1] = key
    conn = np.array(conn, np.int32)
    ne = len(conn)
    par = np.zeros((ne, 8)); par[:, 0] = nip; par[:, 1] = max_iters; par[:, 2] = tol; par[:, 3:6] = np.array(vec)
    fix = np.array([(t, d) for (x, y, z), t in pts.items() if z == 0.0 for d in range(6)], np.int32).reshape(-1, 2)
    loads = []
    for (x, y, z), t in pts.items():
        on_grid = abs(x / bay - round(x / bay)) < 1e-9 and abs(y / bay - round(y / bay)) < 1e-9
        if z > 0 and abs(z / story - round(z / story)) < 1e-9 and on_grid:
            top = round(z / story) == nstory and x == 0.0 and y == 0.0
            loads.append([t, lateral[0] if top else 0.0, lateral[1] if top else 0.0, gravity, 0.0, 0.0, 0.0])
    return ModelSpec(3, 6, np.arange(1, nn + 1, dtype=np.int32), crd, fix, [],
                     [ElementGroup(ELE_FBC3D, np.arange(1, ne + 1, dtype=np.int32), conn, np.ones(ne, np.int32), par)],
                     np.array(loads), uniaxials=[(1, *CONCRETE02_CORE), (2, *CONCRETE02_COVER), (3, *STEEL02)],
                     sections=[rc_section3d(1)])


def soil_structure_block(nx=6, ny=6, nz=6, distort=0.0, seed=0):
    """BASELINE configs[4] in small: a J2 soil block (material 1) with an ElasticIsotropic footing + pier
    (material 2, stiff) embedded at the top centre -- two element batches whose FE_Element (tag) order
    interleaves; loads on the pier top.  ndm 3, ndf 3, stdBrick throughout."""
    spec = brick_block(nx, ny, nz, mat=J2_STEEL, distort=distort, seed=seed)
    g = spec.groups[0]
    e = np.arange(len(g.tags))
    ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
    cx, cy = (nx - 1) / 2.0, (ny - 1) / 2.0
    footing = (ez >= nz - 2) & (np.abs(ex - cx) <= nx / 4.0) & (np.abs(ey - cy) <= ny / 4.0)
    pier = (ez >= nz - 3) & (np.abs(ex - cx) <= 0.5) & (np.abs(ey - cy) <= 0.5)
    struct = footing | pier
    mats = [(1, *J2_STEEL), (2, MAT_ELASTIC, [3.0e6, 0.2, 0.0])]
    groups = [ElementGroup(ELE_BRICK, g.tags[~struct], g.conn[~struct], np.full((~struct).sum(), 1, np.int32), g.par[~struct]),
              ElementGroup(ELE_BRICK, g.tags[struct], g.conn[struct], np.full(struct.sum(), 2, np.int32), g.par[struct])]
    return ModelSpec(3, 3, spec.node_tags, spec.crd, spec.fix, mats, groups, spec.loads)


def tie(spec, u):
    """make a nodal field consistent with the model's `equalDOF`s: constrained dofs take the retained dof's value"""
    if not spec.equal_dofs:
        return u
    ix = {int(t): i for i, t in enumerate(spec.node_tags)}
    for r, c, dofs in spec.equal_dofs:
        u[ix[int(c)], list(dofs)] = u[ix[int(r)], list(dofs)]
    return u


def soil_column_equaldof(ny=8, mat=J2_STEEL, distort=0.15, seed=21):
    """the classic site-response column: one FourNodeQuad wide, base fixed, the left and right node of every
    level tied with `equalDOF` in both dofs (periodic boundary) -- tied dofs sit in the SAME element"""
    spec = quad_plane(1, ny, mat=mat, lx=1.0, ly=float(ny), distort=distort, seed=seed, body=(0.0, -0.02))
    spec.fix = np.array([(t, d) for t in (1, 2) for d in range(2)], np.int32)
    spec.equal_dofs = [(1 + 2 * j, 2 + 2 * j, [0, 1]) for j in range(1, ny + 1)]
    spec.loads = np.array([[1 + 2 * ny, 2.0, -1.0], [2 + 2 * ny, 0.5, 0.25], [1 + 2 * (ny // 2), 1.0, 0.0]])
    return spec


def brick_periodic_equaldof(nx=3, ny=3, nz=2, mat=J2_STEEL, dofs=(0, 2), distort=0.2, seed=22):
    """a stdBrick block whose x = lx face follows its x = 0 face in `dofs` (partial tie, nodes in different
    elements), plus one interior pair tied in all three dofs and one tie onto a fixed node"""
    spec = brick_block(nx, ny, nz, mat=mat, distort=distort, seed=seed, body=(0.0, 0.0, -0.01))
    def nid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k) + 1
    eq = [(nid(0, j, k), nid(nx, j, k), list(dofs)) for k in range(1, nz + 1) for j in range(ny + 1)]
    eq.append((nid(1, 1, 1), nid(nx - 1, ny - 1, nz), [0, 1, 2]))   # far-apart interior / top nodes
    eq.append((nid(1, 0, 0), nid(1, 0, 1), [1]))              # retained dof is fixed: the tied dof is fixed too
    spec.equal_dofs = eq
    return spec


def frame2d_diaphragm_equaldof(nbay=2, nstory=2, ndiv=1, **kw):
    """2D RC frame with every floor's column-line nodes tied to the first one in ux (rigid diaphragm, the usual
    `equalDOF $master $slave 1`): one retained dof with several constrained ones, both ends of a girder on one equation"""
    spec = frame2d(nbay, nstory, ndiv, **kw)
    bay, story = 360.0, 144.0
    eq = []
    for j in range(1, nstory + 1):
        line = [int(spec.node_tags[i]) for i in range(spec.nn)
                if abs(spec.crd[i, 1] - j * story) < 1e-9 and abs(spec.crd[i, 0] / bay - round(spec.crd[i, 0] / bay)) < 1e-9]
        line.sort(key=lambda t: spec.crd[t - 1, 0])
        eq += [(line[0], t, [0]) for t in line[1:]]
    spec.equal_dofs = eq
    return spec


def quad_plane_stress_pressure(nx=8, ny=5, type_=1, pressure=3.0, distort=0.2, seed=31, mat=ELASTIC):
    """FourNodeQuad wall, `PlaneStress` (ElasticIsotropicPlaneStress2D) or PlaneStrain, with the element's surface
    pressure argument set (FourNodeQuad::setPressureLoadAtNodes): par = thickness, type, pressure, rho, b1, b2"""
    spec = quad_plane(nx, ny, mat=mat, lx=float(nx), ly=float(ny), thick=0.4, distort=distort, seed=seed, body=(0.01, -0.03))
    spec.groups[0].par[:, 1] = type_
    spec.groups[0].par[:, 2] = pressure * (1.0 + 0.1 * np.arange(spec.ne))      # a different pressure on every element
    return spec


def have_glue():
    return os.path.exists(GLUE_SO)


def have_metis():
    return os.path.exists(METIS_SO)


def element_graph(spec):
    """Domain::buildEleGraph (domain/domain/Domain.cpp:2410): one vertex per element in ascending tag order,
    an edge between elements that share a node, adjacency sorted (Vertex::addEdge -> ID::insert)"""
    tags = np.concatenate([g.tags for g in spec.groups])
    order = np.argsort(tags, kind="stable")
    conns = [c for g in spec.groups for c in g.conn]
    node2e = {}
    for v, k in enumerate(order):
        for n in conns[k]:
            node2e.setdefault(int(n), []).append(v)
    adj = [set() for _ in order]
    for es in node2e.values():
        for a in es:
            adj[a].update(es)
    xadj, adjncy = [0], []
    for v, s_ in enumerate(adj):
        s_.discard(v)
        adjncy.extend(sorted(s_)); xadj.append(len(adjncy))
    return np.array(xadj, np.int32), np.array(adjncy, np.int32)


def element_graph_fast(spec):
    """the same graph through a sparse product (node-element incidence N: the pattern of N^T N without its diagonal),
    for meshes of 10^6 elements where the set-based walk above takes minutes"""
    import scipy.sparse as sp
    tags = np.concatenate([g.tags for g in spec.groups])
    order = np.argsort(tags, kind="stable")
    conn = np.concatenate([g.conn for g in spec.groups])[order]
    ne, nen = conn.shape
    _, nid = np.unique(conn.ravel(), return_inverse=True)
    N = sp.csr_matrix((np.ones(ne * nen, np.int8), (nid, np.repeat(np.arange(ne), nen))), shape=(nid.max() + 1, ne))
    A = (N.T @ N).tocsr()
    A.setdiag(0); A.eliminate_zeros(); A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32)


def metis_partition(spec, nparts, fast=False):
    """what graph/partitioner/Metis.cpp:320 does for DomainPartitioner: METIS_PartGraphKway (METIS 4 from the
    reference's OTHER/METIS, default options, no weights) on the element graph; part[e] in FE_Element order"""
    L = ctypes.CDLL(METIS_SO)
    xadj, adjncy = element_graph_fast(spec) if fast else element_graph(spec)
    n = ctypes.c_int(len(xadj) - 1); wf = ctypes.c_int(0); nf = ctypes.c_int(0); npart = ctypes.c_int(nparts)
    options = (ctypes.c_int * 5)(0, 0, 0, 0, 0); edgecut = ctypes.c_int(0)
    part = np.zeros(len(xadj) - 1, np.int32)
    if nparts > 1:
        L.METIS_PartGraphKway(ctypes.byref(n), _p(xadj), _p(adjncy), None, None, ctypes.byref(wf), ctypes.byref(nf),
                              ctypes.byref(npart), options, ctypes.byref(edgecut), _p(part))
    return part


def cantilever2d(ndiv=1, nip=5, L=432.0, H=1.0, V=-100.0, max_iters=10, tol=1e-12):
    """the reference's tests/Ex2b.Canti2D.InelasticSection.Push.py cantilever (BASELINE configs[0]) with
    the RC fibre section of north_star (Steel02 + Concrete02): node 1 fixed, forceBeamColumn(s) up to the
    free node, reference load (H lateral, V axial) at the top"""
    nn = ndiv + 1
    crd = np.zeros((nn, 2)); crd[:, 1] = np.linspace(0.0, L, nn)
    conn = np.array([(i + 1, i + 2) for i in range(ndiv)], np.int32)
    par = np.zeros((ndiv, 8)); par[:, 0] = nip; par[:, 1] = max_iters; par[:, 2] = tol
    fix = np.array([(1, 0), (1, 1), (1, 2)], np.int32)
    loads = np.array([[nn, H, V, 0.0]])
    return ModelSpec(2, 3, np.arange(1, nn + 1, dtype=np.int32), crd, fix, [],
                     [ElementGroup(ELE_FBC2D, np.arange(1, ndiv + 1, dtype=np.int32), conn, np.ones(ndiv, np.int32), par)],
                     loads, uniaxials=[(1, *CONCRETE02_CORE), (2, *CONCRETE02_COVER), (3, *STEEL02)],
                     sections=[rc_section(1)])


# tests/Ex2b.Canti2D.InelasticSection.Push.py as written: Steel01 (My, EIcrack, b) on the curvature and an Elastic
# material (EA) on the axial strain, combined by `section Aggregator`
EX2B_MY, EX2B_PHIY, EX2B_B = 130000.0, 0.65e-4, 0.01
EX2B_EA = 57.0 * np.sqrt(4000.0) * (60.0 * 60.0 * 1000.0)
STEEL01_EX2B = (UNI_STEEL01, (EX2B_MY, EX2B_MY / EX2B_PHIY, EX2B_B, 0.0, 55.0, 0.0, 55.0))
ELASTIC_EX2B = (UNI_ELASTIC, (EX2B_EA, 0.0, EX2B_EA))


def cantilever2d_aggregator(ndiv=1, nip=5, L=432.0, H=2000.0, V=0.0, max_iters=10, tol=1e-12):
    """BASELINE configs[0] with the section the script defines: `section Aggregator` of an Elastic axial material and a
    Steel01 moment-curvature material, one forceBeamColumn with 5 Lobatto points, node 1 fixed"""
    spec = cantilever2d(ndiv, nip, L, H, V, max_iters, tol)
    spec.uniaxials = [(3, *ELASTIC_EX2B), (2, *STEEL01_EX2B)]
    spec.sections = [(1, "aggregator", (3, 2))]
    return spec


def ex2b_as_written(glue=False):
    """tests/Ex2b.Canti2D.InelasticSection.Push.py, statement by statement, on the reference's own classes (glue=False)
    or with the integrators' model-facing calls routed to the device (glue=True, oracle/ref_glue.cpp):
    section Aggregator (Elastic P, Steel01 Mz), forceBeamColumn with 5 points, geomTransf Linear; gravity in 10
    LoadControl steps (numberer Plain, system BandGeneral, test NormDispIncr 1e-8 6, algorithm Newton); loadConst -time 0;
    pattern Plain 200 Linear {load 2 Hload 0 0}; DisplacementControl node 2 dof 1 0.001 LCol, test EnergyIncr 1e-8 6;
    50 steps to 5 % drift.  (The script's gravity line is garbled -- "load 2 0 PCol=0" -- and is read as the example
    it was converted from: load 2 0 -PCol 0.)  -> dict of iteration counts, norms, load factors, displacements"""
    LCol, weight = 432.0, 2000.0
    spec = cantilever2d_aggregator(ndiv=1, nip=5, L=LCol, H=0.0, V=-weight)
    R = RefBackend(spec, defer_setup=True, so=GLUE_SO if glue else None)
    out = {}
    if glue: R.setup_glue_loadcontrol(NUMBERER_PLAIN, 2, 0.1, test=0, tol=1e-8, max_iter=6)
    else: R.setup_loadcontrol(NUMBERER_PLAIN, 2, 0.1, test=0, tol=1e-8, max_iter=6)
    rc, it, nm = R.analyze_static(10)
    assert rc == 0, rc
    out["grav_iters"], out["grav_norms"] = it, nm
    out["grav_u"] = R.glue_trial_disp() if glue else R.get_trial_disp()
    R.load_const(0.0)
    R.add_load(2, (weight, 0.0, 0.0))
    if glue: R.setup_glue_dispcontrol(NUMBERER_PLAIN, 2, 2, 0, 0.001 * LCol, test=2, tol=1e-8, max_iter=6)
    else: R.setup_dispcontrol(NUMBERER_PLAIN, 2, 2, 0, 0.001 * LCol, test=2, tol=1e-8, max_iter=6)
    rc, it, nm, lam = R.analyze_static_lam(50)
    assert rc == 0, rc
    out["push_iters"], out["push_norms"], out["push_lam"] = it, nm, lam
    out["push_u"] = R.glue_trial_disp() if glue else R.get_trial_disp()
    if glue:
        out["calls"], out["launches"] = R.glue_counts()
    return out


def ex2b_drive(model, solve, ids, is_dev, tol=1e-10, max_iter=10):
    """the sequence of ex2b_as_written() on any backend's update / form_* surface (oracle on the CPU, device through the
    C-ABI): gravity in 10 LoadControl steps (Newton, NormDispIncr), loadConst -time 0, the lateral pattern, 50
    DisplacementControl steps -> (gravity displacements, push load factors, final displacements)"""
    LCol, weight = 432.0, 2000.0
    if not is_dev: model._u = np.zeros(ids.shape)

    def incr(dU):
        if is_dev:
            model.incr_trial_disp(dU); model.update()
        else:
            model._u[ids >= 0] += dU[ids[ids >= 0]]; model.set_trial_disp(model._u)
    for s in range(10):                               # LoadControl 0.1, Newton, test NormDispIncr
        model.apply_load(0.1 * (s + 1))
        for it in range(max_iter):
            B = model.form_unbalance(); A = model.form_tangent()
            dU = solve(A, B); incr(dU)
            if np.linalg.norm(dU) <= tol: break
        model.commit()
    grav_u = (model.trial_disp() if is_dev else model._u).copy()
    model.load_const(); model.apply_load(0.0)
    model.add_load(2, (weight, 0.0, 0.0))
    hist, lam = disp_control(model, solve, ids[1, 0], 0.001 * LCol, 50, tol, max_iter, is_dev)
    return grav_u, lam, (model.trial_disp() if is_dev else model._u).copy()


def disp_control(model, solve, ctrl_eq, incr, nsteps, tol, max_iter, is_dev):
    """StaticAnalysis with `integrator DisplacementControl node dof incr`, `algorithm Newton`,
    `test NormDispIncr tol max_iter`, driven through any backend's update / form_* surface; the linear
    solves go to `solve(A, b)`.  Follows DisplacementControl::domainChanged / newStep / update
    (analysis/integrator/Static/DisplacementControl.cpp:352-366, 121-207, 210-266) and
    NewtonRaphson::solveCurrentStep.  Returns (norm history per step, lambda per step)."""
    def incr_disp(dU):
        if is_dev:
            model.incr_trial_disp(dU); model.update()
        else:
            u = model._u; ids = model.ids()
            u[ids >= 0] += dU[ids[ids >= 0]]
            model.set_trial_disp(u)
    # domainChanged: phat = unbalance at lambda + 1 ("assumes unbalance at last was 0")
    lam = 0.0
    model.apply_load(lam + 1.0)
    phat = model.form_unbalance().copy()
    hist, lams = [], []
    for _ in range(nsteps):
        # newStep
        A = model.form_tangent()
        dUhat = solve(A, phat)
        dlam = incr / dUhat[ctrl_eq]
        lam += dlam
        incr_disp(dUhat * dlam)
        model.apply_load(lam)
        # NewtonRaphson::solveCurrentStep
        B = model.form_unbalance()
        norms = []
        for it in range(max_iter):
            A = model.form_tangent()
            dUbar = solve(A, B)
            dUhat = solve(A, phat)               # update(): second solve with the reference load
            dL = -dUbar[ctrl_eq] / dUhat[ctrl_eq]
            dU = dUbar + dL * dUhat
            lam += dL
            incr_disp(dU)
            model.apply_load(lam)
            B = model.form_unbalance()
            norms.append(float(np.linalg.norm(dU)))   # the integrator leaves deltaU in X for the test
            if norms[-1] <= tol:
                break
        hist.append(norms); lams.append(lam)
        model.commit()
    return hist, np.array(lams)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _Backend:
    """common numpy-facing surface: setup / ids / csr / update / form_* / commit"""

    def num_eqn(self): ...


class OracleBackend(_Backend):
    def __init__(self, spec: ModelSpec, numberer=NUMBERER_PLAIN, soe=SOE_CSC):
        L = ctypes.CDLL(ORACLE_SO)
        self.L, self.spec = L, spec
        L.orc_model_new.restype = ctypes.c_void_p
        tags = np.ascontiguousarray(spec.node_tags, np.int32)
        crd = np.ascontiguousarray(spec.crd, np.float64)
        self.h = ctypes.c_void_p(L.orc_model_new(spec.ndm, spec.ndf, spec.nn, _p(tags), _p(crd)))
        assert self.h, "node tags must ascend"
        for t, nd in spec.node_ndf.items():
            assert L.orc_set_node_ndf(self.h, int(t), int(nd)) == 0
        for t, d in spec.fix:
            assert L.orc_fix(self.h, int(t), int(d)) == 0
        for r, c, dofs in spec.equal_dofs:
            d = np.ascontiguousarray(dofs, np.int32)
            assert L.orc_equal_dof(self.h, int(r), int(c), len(d), _p(d)) == 0
        for tag, kind, p in spec.materials:
            pp = np.zeros(8); pp[:len(p)] = p
            assert L.orc_add_nd_material(self.h, tag, kind, _p(pp)) == 0
        for tag, kind, p in spec.uniaxials:
            pp = np.zeros(12); pp[:len(p)] = p
            assert L.orc_add_uniaxial(self.h, tag, kind, _p(pp)) == 0
        L.orc_add_fiber_section3d.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        for sec in spec.sections:
            if isinstance(sec[1], str):     # (tag, "aggregator", uniaxial tags for P, Mz)
                mt = np.ascontiguousarray(sec[2], np.int32); codes = np.array([SEC_P, SEC_MZ], np.int32)
                assert L.orc_add_section_aggregator(self.h, sec[0], len(mt), _p(mt), _p(codes)) == 0
                continue
            tag, y, A, mt = sec[:4]
            y, A, mt = np.ascontiguousarray(y, np.float64), np.ascontiguousarray(A, np.float64), np.ascontiguousarray(mt, np.int32)
            if len(sec) == 6:     # 3D: (tag, y, A, mat, z, GJ)
                z = np.ascontiguousarray(sec[4], np.float64)
                assert L.orc_add_fiber_section3d(self.h, tag, len(y), _p(y), _p(z), _p(A), _p(mt), float(sec[5])) == 0
            else:
                assert L.orc_add_fiber_section(self.h, tag, len(y), _p(y), _p(A), _p(mt)) == 0
        for g in spec.groups:
            for i in range(len(g.tags)):
                c = np.ascontiguousarray(g.conn[i], np.int32); pr = np.zeros(16); pr[:len(g.par[i])] = g.par[i]
                rc = L.orc_add_element(self.h, g.kind, int(g.tags[i]), _p(c), int(g.mat[i]), _p(pr))
                assert rc == 0, rc
        if spec.loads is not None:
            for row in spec.loads:
                v = np.ascontiguousarray(row[1:], np.float64)
                assert L.orc_add_load(self.h, int(row[0]), _p(v)) == 0
        if spec.beam_rules:      # (part of the element definition: before the element loads)
            L.orc_set_beam_integration.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
            for t, xi, wt in zip(*spec.beam_rules):
                assert L.orc_set_beam_integration(self.h, int(t), len(xi), _p(np.ascontiguousarray(xi)), _p(np.ascontiguousarray(wt))) == 0
        L.orc_add_beam_uniform_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        for t, wy, wz, wa in spec.beam_loads:
            assert L.orc_add_beam_uniform_load(self.h, int(t), float(wy), float(wz), float(wa)) == 0
        L.orc_add_beam_point_load.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_double] * 4
        for t, py, pz, pn, xl in spec.beam_point_loads:
            assert L.orc_add_beam_point_load(self.h, int(t), float(py), float(pz), float(pn), float(xl)) == 0
        L.orc_add_beam_partial_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        for t, *q in spec.beam_partial_loads:
            assert L.orc_add_beam_partial_load(self.h, int(t), _p(np.array(q, np.float64))) == 0
        # soe 2 / 3 / 4: BandGeneral / ProfileSPD / Umfpack -- the column graph, then the SOE's own storage on top of it
        self.soe = soe
        self.neq = L.orc_setup(self.h, numberer, soe if soe in (0, 1) else 0)
        self.nnz = L.orc_nnz(self.h)
        self.a_size = self.nnz
        if soe in (2, 3):
            L.orc_set_store.restype = ctypes.c_longlong; L.orc_set_store.argtypes = [ctypes.c_void_p, ctypes.c_int]
            self.a_size = int(L.orc_set_store(self.h, soe))
            assert self.a_size >= 0
        self.ne = L.orc_num_ele(self.h)

    def band(self):
        a = np.zeros(2, np.int32); self.L.orc_get_band(self.h, _p(a)); return int(a[0]), int(a[1])

    def profile(self):
        a = np.zeros(self.neq, np.int32); self.L.orc_get_profile(self.h, _p(a)); return a

    def ids(self):
        a = np.zeros((self.spec.nn, self.spec.ndf), np.int32); self.L.orc_get_ids(self.h, _p(a)); return a

    def csr(self):
        ptr = np.zeros(self.neq + 1, np.int32); idx = np.zeros(self.nnz, np.int32)
        self.L.orc_get_csr(self.h, _p(ptr), _p(idx)); return ptr, idx

    def fe_ids(self, stride=24):
        tags = np.zeros(self.ne, np.int32); ids = np.full((self.ne, stride), -9, np.int32)
        self.L.orc_fe_ids(self.h, _p(tags), _p(ids), stride); return tags, ids

    def scatter_map(self, e, nd):
        m = np.zeros(nd * nd, np.int32); self.L.orc_scatter_map(self.h, e, _p(m)); return m.reshape(nd, nd)

    def load_const(self):
        assert self.L.orc_load_const(self.h) == 0

    def add_load(self, node, vals):
        assert self.L.orc_add_load(self.h, int(node), _p(np.ascontiguousarray(vals, np.float64))) == 0

    def set_trial_disp(self, u):
        u = np.ascontiguousarray(u, np.float64); return self.L.orc_set_trial_disp(self.h, _p(u))

    def apply_load(self, lam):
        self.L.orc_apply_load.argtypes = [ctypes.c_void_p, ctypes.c_double]; self.L.orc_apply_load(self.h, lam)

    def form_tangent(self):
        A = np.zeros(self.a_size); self.L.orc_form_tangent(self.h, _p(A)); return A

    def form_unbalance(self):
        B = np.zeros(self.neq); self.L.orc_form_unbalance(self.h, _p(B)); return B

    def ele_tangent(self, e, nd):
        K = np.zeros(nd * nd); self.L.orc_ele_tangent(self.h, e, _p(K)); return K.reshape(nd, nd)

    def ele_resid(self, e, nd):
        R = np.zeros(nd); self.L.orc_ele_resid(self.h, e, _p(R)); return R

    # ---- transient (Newmark, displacement form) ----
    def set_mass(self, tags, mass):
        for t, mv in zip(tags, np.ascontiguousarray(mass, np.float64)):
            assert self.L.orc_set_mass(self.h, int(t), _p(np.ascontiguousarray(mv))) == 0

    def set_rayleigh(self, alphaM, betaK, betaK0, betaKc):
        self.L.orc_set_rayleigh.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
        assert self.L.orc_set_rayleigh(self.h, alphaM, betaK, betaK0, betaKc) == 0

    def set_transient(self, c1, c2, c3):
        self.L.orc_set_transient.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 3
        self.L.orc_set_transient(self.h, c1, c2, c3)

    def newmark_predict(self, a1, a2, a3, a4):
        self.L.orc_newmark_predict.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
        self.L.orc_newmark_predict(self.h, a1, a2, a3, a4)

    def incr_response(self, dU, cu, cv, ca):
        dU = np.ascontiguousarray(dU, np.float64)
        self.L.orc_incr_response.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_double] * 3
        return self.L.orc_incr_response(self.h, _p(dU), cu, cv, ca)

    def set_vel_accel(self, v, a):
        v, a = np.ascontiguousarray(v, np.float64), np.ascontiguousarray(a, np.float64)
        self.L.orc_set_vel_accel(self.h, _p(v), _p(a))

    def vel_accel(self):
        v = np.zeros((self.spec.nn, self.spec.ndf)); a = np.zeros_like(v)
        self.L.orc_get_vel_accel(self.h, _p(v), _p(a)); return v, a

    def update(self):
        return self.set_trial_disp(self.trial_disp())

    def trial_disp(self):
        raise NotImplementedError

    def commit(self):
        return self.L.orc_commit(self.h)

    def revert(self):
        return self.L.orc_revert(self.h)

    def revert_to_start(self):
        return self.L.orc_revert_to_start(self.h)

    def __del__(self):
        try:
            self.L.orc_model_free(self.h)
        except Exception:
            pass


def oracle_nd_path(kind, p, type_, strains, commit):
    L = ctypes.CDLL(ORACLE_SO)
    order = 6 if type_ == ND_3D else 3
    strains = np.ascontiguousarray(strains, np.float64); n = len(strains)
    commit = np.ascontiguousarray(commit, np.int32)
    pp = np.zeros(8); pp[:len(p)] = p
    s = np.zeros((n, order)); t = np.zeros((n, order, order))
    r = L.orc_nd_path(kind, _p(pp), type_, n, _p(strains), _p(commit), _p(s), _p(t))
    assert r == order, r
    return s, t


def have_ref():
    return os.path.exists(REF_SO)


def oracle_uni_path(kind, p, strains, commit):
    L = ctypes.CDLL(ORACLE_SO)
    strains = np.ascontiguousarray(strains, np.float64); commit = np.ascontiguousarray(commit, np.int32)
    pp = np.zeros(12); pp[:len(p)] = p
    s = np.zeros(len(strains)); t = np.zeros(len(strains))
    assert L.orc_uni_path(kind, _p(pp), len(strains), _p(strains), _p(commit), _p(s), _p(t)) == 0
    return s, t


def ref_uni_path(kind, p, strains, commit):
    L = ctypes.CDLL(REF_SO)
    strains = np.ascontiguousarray(strains, np.float64); commit = np.ascontiguousarray(commit, np.int32)
    pp = np.zeros(12); pp[:len(p)] = p
    s = np.zeros(len(strains)); t = np.zeros(len(strains))
    assert L.ref_uni_path(kind, _p(pp), len(strains), _p(strains), _p(commit), _p(s), _p(t)) == 0
    return s, t


def ref_nd_path(kind, p, type_, strains, commit):
    L = ctypes.CDLL(REF_SO)
    order = 6 if type_ == ND_3D else 3
    strains = np.ascontiguousarray(strains, np.float64); n = len(strains)
    commit = np.ascontiguousarray(commit, np.int32)
    pp = np.zeros(8); pp[:len(p)] = p
    s = np.zeros((n, order)); t = np.zeros((n, order, order))
    name = {ND_3D: b"ThreeDimensional", ND_PLANE_STRAIN: b"PlaneStrain", ND_PLANE_STRESS: b"PlaneStress"}[type_]
    r = L.ref_nd_path(kind, _p(pp), name, n, _p(strains), _p(commit), _p(s), _p(t))
    assert r == order, r
    return s, t


class RefBackend(_Backend):
    """The reference's own Domain / AnalysisModel / LinearSOE, through oracle/ref_harness.cpp."""

    def __init__(self, spec: ModelSpec, numberer=NUMBERER_PLAIN, soe=SOE_CSC, dlambda=1.0,
                 test=0, tol=1e-8, max_iter=20, defer_setup=False, so=None, handler=0):
        L = ctypes.CDLL(so or REF_SO)
        self.L, self.spec = L, spec
        L.ref_model_new.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(L.ref_model_new(spec.ndm, spec.ndf))
        self.tags = np.ascontiguousarray(spec.node_tags, np.int32)
        for t, x in zip(spec.node_tags, spec.crd):
            xx = np.zeros(3); xx[:spec.ndm] = x
            if int(t) in spec.node_ndf:
                assert L.ref_add_node_ndf(self.h, int(t), _p(xx), int(spec.node_ndf[int(t)])) == 0
            else:
                assert L.ref_add_node(self.h, int(t), _p(xx)) == 0
        for t, d in spec.fix:
            assert L.ref_fix(self.h, int(t), int(d)) == 0
        for r, c, dofs in spec.equal_dofs:
            d = np.ascontiguousarray(dofs, np.int32)
            assert L.ref_equal_dof(self.h, int(r), int(c), len(d), _p(d)) == 0
        for tag, kind, p in spec.materials:
            pp = np.zeros(8); pp[:len(p)] = p
            assert L.ref_add_nd_material(self.h, tag, kind, _p(pp)) == 0
        L.ref_add_quad.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        L.ref_add_force_beam2d_t.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_double, ctypes.c_int]
        for tag, kind, p in spec.uniaxials:
            pp = np.zeros(12); pp[:len(p)] = p
            assert L.ref_add_uniaxial(self.h, tag, kind, _p(pp)) == 0
        L.ref_add_fiber_section3d.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        L.ref_add_force_beam3d_t.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
        for sec in spec.sections:
            if isinstance(sec[1], str):     # (tag, "aggregator", uniaxial tags for P, Mz)
                mt = np.ascontiguousarray(sec[2], np.int32); codes = np.array([SEC_P, SEC_MZ], np.int32)
                assert L.ref_add_section_aggregator(self.h, sec[0], len(mt), _p(mt), _p(codes)) == 0
                continue
            tag, y, A, mt = sec[:4]
            y, A, mt = np.ascontiguousarray(y, np.float64), np.ascontiguousarray(A, np.float64), np.ascontiguousarray(mt, np.int32)
            if len(sec) == 6:
                z = np.ascontiguousarray(sec[4], np.float64)
                assert L.ref_add_fiber_section3d(self.h, tag, len(y), _p(y), _p(z), _p(A), _p(mt), float(sec[5])) == 0
            else:
                assert L.ref_add_fiber_section(self.h, tag, len(y), _p(y), _p(A), _p(mt)) == 0
        self.ele_tags = []
        for g in spec.groups:
            for i in range(len(g.tags)):
                c = np.ascontiguousarray(g.conn[i], np.int32)
                if g.kind == ELE_BRICK:
                    b = np.ascontiguousarray(g.par[i, :3], np.float64)
                    assert L.ref_add_brick(self.h, int(g.tags[i]), _p(c), int(g.mat[i]), _p(b)) == 0
                elif g.kind == ELE_FBC3D:
                    L.ref_set_beam_rho.argtypes = [ctypes.c_void_p, ctypes.c_double]
                    L.ref_set_beam_rho(self.h, float(g.par[i, 7]))
                    o6 = np.zeros(6)
                    if g.par.shape[1] >= 14: o6[:] = g.par[i, 8:14]
                    L.ref_set_beam_offsets(self.h, _p(o6))
                    vx = np.ascontiguousarray(g.par[i, 3:6], np.float64)
                    assert L.ref_add_force_beam3d_t(self.h, int(g.tags[i]), _p(c), int(g.mat[i]), int(g.par[i, 0]),
                                                    int(g.par[i, 1]), float(g.par[i, 2]), _p(vx), int(g.par[i, 6]) + 16 * spec.beam_integration) == 0
                elif g.kind == ELE_FBC2D:
                    L.ref_set_beam_rho.argtypes = [ctypes.c_void_p, ctypes.c_double]
                    L.ref_set_beam_rho(self.h, float(g.par[i, 4]))
                    o6 = np.zeros(6)
                    if g.par.shape[1] >= 9: o6[0:2] = g.par[i, 5:7]; o6[3:5] = g.par[i, 7:9]
                    L.ref_set_beam_offsets(self.h, _p(o6))
                    assert L.ref_add_force_beam2d_t(self.h, int(g.tags[i]), _p(c), int(g.mat[i]), int(g.par[i, 0]),
                                                    int(g.par[i, 1]), float(g.par[i, 2]), int(g.par[i, 3]) + 16 * spec.beam_integration) == 0
                else:
                    b = np.ascontiguousarray(g.par[i, 4:6], np.float64)
                    assert L.ref_add_quad(self.h, int(g.tags[i]), _p(c), int(g.mat[i]), float(g.par[i, 0]),
                                          int(g.par[i, 1]), float(g.par[i, 2]), float(g.par[i, 3]), _p(b)) == 0
                self.ele_tags.append(int(g.tags[i]))
        self.ele_tags.sort()
        if spec.loads is not None:
            for row in spec.loads:
                v = np.ascontiguousarray(row[1:], np.float64)
                assert L.ref_add_load(self.h, int(row[0]), _p(v)) == 0
        if spec.beam_partial_loads:
            L.ref_add_beam_partial_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            for t, *q in spec.beam_partial_loads:
                assert L.ref_add_beam_partial_load(self.h, int(t), _p(np.array(q, np.float64))) == 0
        if spec.beam_point_loads:
            L.ref_add_beam_point_load.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_double] * 4
            for t, py, pz, pn, xl in spec.beam_point_loads:
                assert L.ref_add_beam_point_load(self.h, int(t), float(py), float(pz), float(pn), float(xl)) == 0
        if spec.beam_loads:
            L.ref_add_beam_uniform_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
            for t, wy, wz, wa in spec.beam_loads:
                assert L.ref_add_beam_uniform_load(self.h, int(t), float(wy), float(wz), float(wa)) == 0
        self.ne = len(self.ele_tags)
        self.max_iter = max_iter
        L.ref_set_handler(self.h, int(handler))      # `constraints Plain` (0) | `constraints Transformation` (1)
        if defer_setup:       # the caller picks the integrator (setup_transient)
            return
        L.ref_setup.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                ctypes.c_double, ctypes.c_int]
        self.neq = L.ref_setup(self.h, numberer, soe, dlambda, test, tol, max_iter)
        assert self.neq >= 0, self.neq
        self.nnz = L.ref_nnz(self.h)
        self.ne = len(self.ele_tags)
        self.max_iter = max_iter

    def ids(self):
        a = np.zeros((self.spec.nn, self.spec.ndf), np.int32)
        row = np.zeros(self.spec.ndf, np.int32)
        for i, t in enumerate(self.spec.node_tags):
            row[:] = -1                                        # (a node with fewer dofs fills its first entries only)
            self.L.ref_node_ids(self.h, int(t), _p(row)); a[i] = row
        return a

    def csr(self):
        ptr = np.zeros(self.neq + 1, np.int32); idx = np.zeros(self.nnz, np.int32)
        self.L.ref_get_csr(self.h, _p(ptr), _p(idx)); return ptr, idx

    def fe_ids(self, stride=24):
        tags = np.zeros(self.ne, np.int32); ids = np.full((self.ne, stride), -9, np.int32)
        self.L.ref_fe_ids(self.h, _p(tags), _p(ids), stride); return tags, ids

    def set_trial_disp(self, u):
        u = np.ascontiguousarray(u, np.float64)
        return self.L.ref_set_trial_disp(self.h, self.spec.nn, _p(self.tags), _p(u))

    def get_trial_disp(self):
        u = np.zeros((self.spec.nn, self.spec.ndf))
        self.L.ref_get_trial_disp(self.h, self.spec.nn, _p(self.tags), _p(u)); return u

    def apply_load(self, lam):
        self.L.ref_apply_load.argtypes = [ctypes.c_void_p, ctypes.c_double]; self.L.ref_apply_load(self.h, lam)

    def form_tangent(self):
        A = np.zeros(self.nnz); self.L.ref_form_tangent(self.h, _p(A)); return A

    def form_unbalance(self):
        B = np.zeros(self.neq); self.L.ref_form_unbalance(self.h, _p(B)); return B

    def ele_tangent(self, e, nd):
        K = np.zeros(nd * nd); self.L.ref_ele_tangent(self.h, self.ele_tags[e], _p(K)); return K.reshape(nd, nd)

    def ele_resid(self, e, nd):
        R = np.zeros(nd); self.L.ref_ele_resid(self.h, self.ele_tags[e], _p(R)); return R

    def commit(self):
        return self.L.ref_commit(self.h)

    def revert(self):
        return self.L.ref_revert(self.h)

    def revert_to_start(self):
        return self.L.ref_revert_to_start(self.h)

    # ---- transient: the reference's own Newmark ----
    def set_mass(self, tags, mass):
        for t, mv in zip(tags, np.ascontiguousarray(mass, np.float64)):
            assert self.L.ref_set_mass(self.h, int(t), _p(np.ascontiguousarray(mv))) == 0

    # ---- the reference's own analysis loop on top of the device path (oracle/ref_glue.cpp, so=GLUE_SO) ----
    def setup_glue_loadcontrol(self, numberer, soe, dlambda, test=0, tol=1e-8, max_iter=20, device=0):
        self.L.glue_setup_loadcontrol.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                                  ctypes.c_double, ctypes.c_int, ctypes.c_int]
        self.neq = self.L.glue_setup_loadcontrol(self.h, numberer, soe, dlambda, test, tol, max_iter, device)
        if self.neq < 0:
            self.L.glue_last_error.restype = ctypes.c_char_p; self.L.glue_last_error.argtypes = [ctypes.c_void_p]
            raise RuntimeError(f"glue set-up failed ({self.neq}): {self.L.glue_last_error(self.h).decode()}")
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def setup_glue_newmark(self, numberer, soe, gamma, beta, test=0, tol=1e-8, max_iter=20, device=0):
        self.L.glue_setup_newmark.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        self.neq = self.L.glue_setup_newmark(self.h, numberer, soe, gamma, beta, test, tol, max_iter, device)
        if self.neq < 0:
            self.L.glue_last_error.restype = ctypes.c_char_p; self.L.glue_last_error.argtypes = [ctypes.c_void_p]
            raise RuntimeError(f"glue set-up failed ({self.neq}): {self.L.glue_last_error(self.h).decode()}")
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def setup_glue_dispcontrol(self, numberer, soe, node, dof, incr, test=0, tol=1e-8, max_iter=20, device=0):
        self.L.glue_setup_dispcontrol.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        self.neq = self.L.glue_setup_dispcontrol(self.h, numberer, soe, node, dof, incr, test, tol, max_iter, device)
        if self.neq < 0:
            self.L.glue_last_error.restype = ctypes.c_char_p; self.L.glue_last_error.argtypes = [ctypes.c_void_p]
            raise RuntimeError(f"glue set-up failed ({self.neq}): {self.L.glue_last_error(self.h).decode()}")
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def store_tangent(self, kind):
        """the reference's own BandGenLinSOE (2) / ProfileSPDLinSOE (3) sized and filled from this analysis:
        -> (layout, A): layout = (numSubD, numSuperD) or iDiagLoc"""
        self.L.ref_store_tangent.restype = ctypes.c_longlong
        self.L.ref_store_tangent.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lay = np.zeros(max(self.neq, 2), np.int32)
        n = int(self.L.ref_store_tangent(self.h, kind, _p(lay), None))
        assert n >= 0
        A = np.zeros(n)
        assert int(self.L.ref_store_tangent(self.h, kind, _p(lay), _p(A))) == n
        return ((int(lay[0]), int(lay[1])) if kind == 2 else lay[:self.neq].copy()), A

    def glue_free(self):
        """frees the device model behind a glued reference model (oracle/ref_glue.cpp glue_destroy)"""
        if hasattr(self.L, "glue_destroy"):
            self.L.glue_destroy.argtypes = [ctypes.c_void_p]; self.L.glue_destroy.restype = None
            self.L.glue_destroy(self.h)

    def __del__(self):
        try:
            self.glue_free()
        except Exception:
            pass

    def glue_counts(self):
        c = (ctypes.c_long * 4)(); self.L.glue_call_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        self.L.glue_call_counts(self.h, c)
        self.L.glue_launch_count.restype = ctypes.c_longlong; self.L.glue_launch_count.argtypes = [ctypes.c_void_p]
        return list(c), int(self.L.glue_launch_count(self.h))

    def glue_trial_disp(self):
        u = np.zeros((self.spec.nn, self.spec.ndf)); self.L.glue_get_trial_disp.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert self.L.glue_get_trial_disp(self.h, _p(u)) >= 0
        return u

    def set_rayleigh(self, alphaM, betaK, betaK0, betaKc):
        self.L.ref_set_rayleigh.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
        assert self.L.ref_set_rayleigh(self.h, alphaM, betaK, betaK0, betaKc) == 0

    def setup_transient(self, numberer, soe, gamma, beta, test=0, tol=1e-8, max_iter=20):
        self.L.ref_setup_transient.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_int, ctypes.c_double, ctypes.c_int]
        self.neq = self.L.ref_setup_transient(self.h, numberer, soe, gamma, beta, test, tol, max_iter)
        assert self.neq >= 0
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def new_step(self, dt):
        self.L.ref_transient_new_step.argtypes = [ctypes.c_void_p, ctypes.c_double]
        return self.L.ref_transient_new_step(self.h, dt)

    def transient_update(self, dU):
        dU = np.ascontiguousarray(dU, np.float64)
        return self.L.ref_transient_update(self.h, _p(dU))

    def vel_accel(self):
        v = np.zeros((self.spec.nn, self.spec.ndf)); a = np.zeros_like(v)
        self.L.ref_get_vel_accel(self.h, self.spec.nn, _p(self.tags), _p(v), _p(a)); return v, a

    def analyze_transient(self, nsteps, dt):
        self.L.ref_analyze_transient.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        iters = np.zeros(nsteps, np.int32); norms = np.zeros((nsteps, self.max_iter))
        rc = self.L.ref_analyze_transient(self.h, nsteps, dt, _p(iters), _p(norms), self.max_iter)
        return rc, iters, norms

    def analyze_static(self, nsteps):
        iters = np.zeros(nsteps, np.int32); norms = np.zeros((nsteps, self.max_iter))
        rc = self.L.ref_analyze_static(self.h, nsteps, _p(iters), _p(norms), self.max_iter)
        return rc, iters, norms

    # ---- `analysis Static` set up after the model (integrator LoadControl), loadConst, a further load pattern ----
    def setup_loadcontrol(self, numberer, soe, dlambda, test=0, tol=1e-8, max_iter=20):
        self.L.ref_setup.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        self.neq = self.L.ref_setup(self.h, numberer, soe, dlambda, test, tol, max_iter)
        assert self.neq >= 0, self.neq
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def load_const(self, time=0.0):
        """loadConst -time t; the next add_load opens a new `pattern Plain`"""
        self.L.ref_load_const.argtypes = [ctypes.c_void_p, ctypes.c_double]
        assert self.L.ref_load_const(self.h, float(time)) == 0

    def add_load(self, node, vals):
        assert self.L.ref_add_load(self.h, int(node), _p(np.ascontiguousarray(vals, np.float64))) == 0

    # ---- the reference's own DisplacementControl ----
    def setup_dispcontrol(self, numberer, soe, node, dof, incr, test=0, tol=1e-8, max_iter=20):
        self.L.ref_setup_dispcontrol.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        self.neq = self.L.ref_setup_dispcontrol(self.h, numberer, soe, node, dof, incr, test, tol, max_iter)
        assert self.neq >= 0, self.neq
        self.nnz = self.L.ref_nnz(self.h); self.max_iter = max_iter

    def analyze_static_lam(self, nsteps):
        iters = np.zeros(nsteps, np.int32); norms = np.zeros((nsteps, self.max_iter)); lam = np.zeros(nsteps)
        rc = self.L.ref_analyze_static_lam(self.h, nsteps, _p(iters), _p(norms), self.max_iter, _p(lam))
        return rc, iters, norms, lam
