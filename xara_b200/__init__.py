"""xara_b200 -- B200-native element state determination and global assembly.

Python face of the C-ABI in ``include/xara_b200.h`` (ctypes, no torch types cross the
boundary).  It mirrors the reference's analysis-side vocabulary for the hot path:

==============================  =====================================================
reference (OpenSeesRT / xara)   here
==============================  =====================================================
``model basic -ndm -ndf``       ``DeviceModel(ndm, ndf)``
``node`` / ``fix``              ``add_nodes`` / ``fix``
``nDMaterial``                  ``nd_material``
``element stdBrick | quad``     ``add_elements``
``load`` in ``pattern Plain``   ``add_nodal_loads``
``numberer`` + ``system``       ``setup(numberer, soe)``   (domainChanged)
``Domain::update``              ``update``
``formTangent/formUnbalance``   ``form_tangent`` / ``form_unbalance``
``Domain::commit``              ``commit``
==============================  =====================================================

There is no CPU fallback: if ``libxara_b200.so`` is missing the import fails, and every
device call raises ``XaraB200Error`` when no sm_100a GPU is present.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

__all__ = ["DeviceModel", "XaraB200Error", "lib", "LIB_PATH", "device_count", "comm_unique_id", "exchange_local",
           "MAT_ELASTIC_ISOTROPIC", "MAT_J2PLASTICITY", "ELE_STDBRICK", "ELE_FOURNODEQUAD", "ELE_FORCEBEAMCOLUMN2D", "ELE_FORCEBEAMCOLUMN3D", "UNI_STEEL02", "UNI_CONCRETE02", "UNI_STEEL01", "UNI_ELASTIC", "UNI_CONCRETE01", "UNI_ELASTICPP",
           "NUMBERER_PLAIN", "NUMBERER_RCM", "SOE_SPARSE_GEN_COL", "SOE_SPARSE_GEN_ROW", "SOE_BAND_GEN", "SOE_PROFILE_SPD", "SOE_UMFPACK_GEN"]

MAT_ELASTIC_ISOTROPIC, MAT_J2PLASTICITY = 0, 1
ELE_STDBRICK, ELE_FOURNODEQUAD, ELE_FORCEBEAMCOLUMN2D, ELE_FORCEBEAMCOLUMN3D = 0, 1, 2, 3
UNI_STEEL02, UNI_CONCRETE02, UNI_STEEL01, UNI_ELASTIC, UNI_CONCRETE01, UNI_ELASTICPP = 0, 1, 2, 3, 4, 5
NUMBERER_PLAIN, NUMBERER_RCM = 0, 1
SOE_SPARSE_GEN_COL, SOE_SPARSE_GEN_ROW, SOE_BAND_GEN, SOE_PROFILE_SPD, SOE_UMFPACK_GEN = 0, 1, 2, 3, 4

# XARA_B200_LIB: an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("XARA_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libxara_b200.so")


class XaraB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(xara_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
    sig = {
        "xb_version": (ctypes.c_char_p, []),
        "xb_last_error": (ctypes.c_char_p, []),
        "xb_device_count": (i32, []),
        "xb_model_create": (vp, [i32, i32]),
        "xb_model_destroy": (None, [vp]),
        "xb_add_nodes": (i32, [vp, i32, vp, vp]),
        "xb_set_node_ndf": (i32, [vp, i32, vp, i32]),
        "xb_add_sp": (i32, [vp, i32, vp, vp]),
        "xb_add_equal_dof": (i32, [vp, i32, i32, i32, vp]),
        "xb_add_nd_material": (i32, [vp, i32, i32, vp, i32]),
        "xb_add_uniaxial_material": (i32, [vp, i32, i32, vp, i32]),
        "xb_add_fiber_section": (i32, [vp, i32, i32, vp, vp, vp]),
        "xb_add_fiber_section3d": (i32, [vp, i32, i32, vp, vp, vp, vp, f64]),
        "xb_add_section_aggregator": (i32, [vp, i32, i32, vp, vp]),
        "xb_load_const": (i32, [vp]),
        "xb_set_nodal_loads": (i32, [vp, i32, vp, vp]),
        "xb_add_elements": (i32, [vp, i32, i32, vp, vp, vp, vp, i32]),
        "xb_add_beam_uniform_loads": (i32, [vp, i32, vp, vp]),
        "xb_add_beam_point_loads": (i32, [vp, i32, vp, vp]),
        "xb_add_beam_partial_loads": (i32, [vp, i32, vp, vp]),
        "xb_set_beam_integration": (i32, [vp, i32, vp, i32, vp, vp]),
        "xb_add_nodal_loads": (i32, [vp, i32, vp, vp]),
        "xb_set_nodal_mass": (i32, [vp, i32, vp, vp]),
        "xb_set_rayleigh_alpha_m": (i32, [vp, f64]),
        "xb_set_rayleigh": (i32, [vp, f64, f64, f64, f64]),
        "xb_set_transient_factors": (i32, [vp, f64, f64, f64]),
        "xb_newmark_predict": (i32, [vp, f64, f64, f64, f64]),
        "xb_incr_trial_response": (i32, [vp, vp, f64, f64, f64]),
        "xb_set_trial_vel_accel": (i32, [vp, vp, vp]),
        "xb_get_trial_vel_accel": (i32, [vp, vp, vp]),
        "xb_setup": (i32, [vp, i32, i32]),
        "xb_setup_partitioned": (i32, [vp, i32, i32, i32, i32, vp]),
        "xb_num_rows": (i32, [vp]),
        "xb_get_row_eqns": (i32, [vp, vp]),
        "xb_get_partition": (i32, [vp, vp]),
        "xb_num_peers": (i32, [vp]),
        "xb_get_peer": (i32, [vp, i32, vp, vp]),
        "xb_comm_unique_id": (i32, [vp]),
        "xb_comm_init": (i32, [vp, vp]),
        "xb_exchange": (i32, [vp, i32]),
        "xb_exchange_local": (i32, [vp, i32, i32]),
        "xb_num_nodes": (i32, [vp]),
        "xb_num_elements": (i64, [vp]),
        "xb_num_gauss_points": (i64, [vp]),
        "xb_num_eqn": (i32, [vp]),
        "xb_nnz": (i64, [vp]),
        "xb_a_size": (i64, [vp]),
        "xb_get_band": (i32, [vp, vp, vp]),
        "xb_get_profile": (i32, [vp, vp]),
        "xb_get_node_tags": (i32, [vp, vp]),
        "xb_get_ids": (i32, [vp, vp]),
        "xb_get_element_tags": (i32, [vp, vp]),
        "xb_get_pattern": (i32, [vp, vp, vp]),
        "xb_get_scatter_map": (i32, [vp, i64, i64, vp]),
        "xb_device_init": (i32, [vp, i32, vp]),
        "xb_set_trial_disp": (i32, [vp, vp]),
        "xb_incr_trial_disp": (i32, [vp, vp]),
        "xb_get_trial_disp": (i32, [vp, vp]),
        "xb_update": (i32, [vp]),
        "xb_apply_load": (i32, [vp, f64]),
        "xb_set_load_factor": (i32, [vp, f64]),
        "xb_form_tangent": (i32, [vp, vp]),
        "xb_form_unbalance": (i32, [vp, vp]),
        "xb_form_element_tangents": (i32, [vp]),
        "xb_assemble_tangent": (i32, [vp, vp]),
        "xb_form_element_resids": (i32, [vp]),
        "xb_assemble_unbalance": (i32, [vp, vp]),
        "xb_commit": (i32, [vp]),
        "xb_revert_to_last_commit": (i32, [vp]),
        "xb_revert_to_start": (i32, [vp]),
        "xb_synchronize": (i32, [vp]),
        "xb_device_A": (vp, [vp]),
        "xb_device_B": (vp, [vp]),
        "xb_device_trial_disp": (vp, [vp]),
        "xb_get_element_tangent": (i32, [vp, i64, vp]),
        "xb_get_element_resid": (i32, [vp, i64, vp]),
        "xb_get_gp_response": (i32, [vp, i64, i32, vp, vp]),
        "xb_launch_count": (i64, [vp]),
        "xb_set_option": (i32, [vp, ctypes.c_char_p, i32]),
        "xb_algorithmic_bytes": (i64, [vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L, tuple(sig)


lib, EXPORTS = _load()


def device_count() -> int:
    return lib.xb_device_count()


def comm_unique_id() -> bytes:
    """ncclGetUniqueId (128 bytes); broadcast it to every rank and hand it to DeviceModel.comm_init"""
    buf = ctypes.create_string_buffer(128)
    rc = lib.xb_comm_unique_id(ctypes.addressof(buf))
    if rc < 0:
        raise XaraB200Error(f"[{rc}] {lib.xb_last_error().decode()}")
    return buf.raw


def exchange_local(models, which):
    """interface exchange between the ranks of one partition held in THIS process (device copies)"""
    arr = (ctypes.c_void_p * len(models))(*[m._h for m in models])
    rc = lib.xb_exchange_local(ctypes.addressof(arr), len(models), which)
    if rc < 0:
        raise XaraB200Error(f"[{rc}] {lib.xb_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class DeviceModel:
    """A Domain + AnalysisModel + LinearSOE whose hot path runs on one B200."""

    def __init__(self, ndm: int, ndf: int):
        self._h = lib.xb_model_create(ndm, ndf)
        if not self._h:
            raise XaraB200Error(lib.xb_last_error().decode())
        self.ndm, self.ndf = ndm, ndf
        self.neq = None

    def _ck(self, rc):
        if rc < 0:
            raise XaraB200Error(f"[{rc}] {lib.xb_last_error().decode()}")
        return rc

    def close(self):
        if getattr(self, "_h", None):
            lib.xb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model commands ----
    def add_nodes(self, tags, crd):
        tags, crd = _i32(tags), _f64(crd)
        assert crd.shape == (len(tags), self.ndm)
        self._ck(lib.xb_add_nodes(self._h, len(tags), _ptr(tags), _ptr(crd)))

    def fix(self, node_tags, dofs):
        node_tags, dofs = _i32(node_tags), _i32(dofs)
        self._ck(lib.xb_add_sp(self._h, len(node_tags), _ptr(node_tags), _ptr(dofs)))

    def equal_dof(self, retained_tag, constrained_tag, dofs):
        dofs = np.ascontiguousarray(dofs, np.int32)
        self._ck(lib.xb_add_equal_dof(self._h, int(retained_tag), int(constrained_tag), len(dofs), _ptr(dofs)))

    def nd_material(self, tag, kind, par):
        par = _f64(par)
        self._ck(lib.xb_add_nd_material(self._h, tag, kind, _ptr(par), len(par)))

    def uniaxial_material(self, tag, kind, par):
        par = _f64(par)
        self._ck(lib.xb_add_uniaxial_material(self._h, tag, kind, _ptr(par), len(par)))

    def fiber_section(self, tag, y, A, mat_tags):
        y, A, mat_tags = _f64(y), _f64(A), _i32(mat_tags)
        self._ck(lib.xb_add_fiber_section(self._h, tag, len(y), _ptr(y), _ptr(A), _ptr(mat_tags)))

    def section_aggregator(self, tag, mat_tags, codes=(2, 1)):
        """section Aggregator tag mat P mat Mz (codes of SectionForceDeformation.h: 2 = P, 1 = Mz)"""
        mat_tags, codes = _i32(mat_tags), _i32(codes)
        self._ck(lib.xb_add_section_aggregator(self._h, tag, len(mat_tags), _ptr(mat_tags), _ptr(codes)))

    def fiber_section3d(self, tag, y, z, A, mat_tags, GJ):
        y, z, A, mat_tags = _f64(y), _f64(z), _f64(A), _i32(mat_tags)
        self._ck(lib.xb_add_fiber_section3d(self._h, tag, len(y), _ptr(y), _ptr(z), _ptr(A), _ptr(mat_tags), float(GJ)))

    def add_elements(self, kind, tags, conn, mat_tags, par):
        tags, conn, mat_tags, par = _i32(tags), _i32(conn), _i32(mat_tags), _f64(par)
        assert par.ndim == 2 and len(par) == len(tags)
        self._ck(lib.xb_add_elements(self._h, kind, len(tags), _ptr(tags), _ptr(conn), _ptr(mat_tags),
                                     _ptr(par), par.shape[1]))

    def set_beam_integration(self, ele_tags, xi, wt):
        """section locations / weights [n][nip] (fractions of L) of a beam integration other than Lobatto"""
        ele_tags, xi, wt = _i32(ele_tags), _f64(xi), _f64(wt)
        self._ck(lib.xb_set_beam_integration(self._h, len(ele_tags), _ptr(ele_tags), xi.shape[1], _ptr(xi), _ptr(wt)))

    def add_beam_point_loads(self, ele_tags, p):
        """eleLoad -beamPoint: p [n][4] = Py, Pz, N, xL per element"""
        ele_tags, p = _i32(ele_tags), _f64(p)
        self._ck(lib.xb_add_beam_point_loads(self._h, len(ele_tags), _ptr(ele_tags), _ptr(p)))

    def add_beam_partial_loads(self, ele_tags, p):
        """`eleLoad -beamUniform` over part of an element: p [n][8] = wya, wyb, waa, wab, aOverL, bOverL, wza, wzb (3D)"""
        ele_tags, p = _i32(ele_tags), _f64(p)
        assert p.ndim == 2 and p.shape[1] == 8
        self._ck(lib.xb_add_beam_partial_loads(self._h, len(ele_tags), _ptr(ele_tags), _ptr(p)))

    def add_beam_uniform_loads(self, ele_tags, w):
        """`eleLoad -beamUniform`: w [n][3] = wy, wz, wa per element (Linear pattern)"""
        ele_tags, w = _i32(ele_tags), _f64(w)
        assert w.ndim == 2 and w.shape[1] == 3 and len(w) == len(ele_tags)
        self._ck(lib.xb_add_beam_uniform_loads(self._h, len(ele_tags), _ptr(ele_tags), _ptr(w)))

    def set_node_ndf(self, node_tags, ndf):
        """nodes created under another `model -ndf`: they carry `ndf` (< the model's) dofs"""
        node_tags = _i32(node_tags)
        self._ck(lib.xb_set_node_ndf(self._h, len(node_tags), _ptr(node_tags), int(ndf)))

    def add_nodal_loads(self, node_tags, values):
        node_tags, values = _i32(node_tags), _f64(values)
        self._ck(lib.xb_add_nodal_loads(self._h, len(node_tags), _ptr(node_tags), _ptr(values)))

    @classmethod
    def from_spec(cls, spec, numberer=NUMBERER_PLAIN, soe=SOE_SPARSE_GEN_COL, nparts=1, rank=0, part=None, setup=True, options=None):
        """Build from a tests/modelspec.py ModelSpec (duck-typed); setup=False leaves xb_setup to the caller
        (e.g. to add nodal masses first); options: {name: value} for xb_set_option before the set-up."""
        m = cls(spec.ndm, spec.ndf)
        for k, v in (options or {}).items():
            m.set_option(k, v)
        m.add_nodes(spec.node_tags, spec.crd)
        nn_ = getattr(spec, "node_ndf", {})
        for nd in sorted(set(nn_.values())):
            m.set_node_ndf([t for t, v in nn_.items() if v == nd], nd)
        if len(spec.fix):
            m.fix(spec.fix[:, 0], spec.fix[:, 1])
        for r, c, dofs in getattr(spec, "equal_dofs", []):
            m.equal_dof(r, c, dofs)
        for tag, kind, p in spec.materials:
            m.nd_material(tag, kind, p)
        for tag, kind, p in getattr(spec, "uniaxials", []):
            m.uniaxial_material(tag, kind, p)
        for sec in getattr(spec, "sections", []):
            if isinstance(sec[1], str):      # (tag, "aggregator", uniaxial tags for P, Mz)
                m.section_aggregator(sec[0], sec[2])
            elif len(sec) == 6:      # 3D: (tag, y, A, mat, z, GJ)
                m.fiber_section3d(sec[0], sec[1], sec[4], sec[2], sec[3], sec[5])
            else:
                m.fiber_section(*sec)
        for g in spec.groups:
            m.add_elements(g.kind, g.tags, g.conn, g.mat, g.par)
        if spec.loads is not None and len(spec.loads):
            m.add_nodal_loads(spec.loads[:, 0].astype(np.int32), spec.loads[:, 1:])
        br = getattr(spec, "beam_rules", None)
        if br:
            m.set_beam_integration(*br)
        bp = getattr(spec, "beam_point_loads", [])
        if bp:
            m.add_beam_point_loads([t for t, *_ in bp], np.array([q for _, *q in bp], np.float64))
        bq = getattr(spec, "beam_partial_loads", [])
        if bq:
            m.add_beam_partial_loads([t for t, *_ in bq], np.array([(list(q) + [0.0, 0.0])[:8] for _, *q in bq], np.float64))
        bl = getattr(spec, "beam_loads", [])
        if bl:
            m.add_beam_uniform_loads([t for t, *_ in bl], np.array([w for _, *w in bl], np.float64))
        if setup:
            m.setup(numberer, soe, nparts, rank, part)
        return m

    # ---- analysis set-up (host) ----
    def setup(self, numberer=NUMBERER_PLAIN, soe=SOE_SPARSE_GEN_COL, nparts=1, rank=0, part=None):
        """domainChanged().  nparts > 1: this process is `rank` of a partitioned run (every rank is
        given the whole model); part = None -> built-in coordinate bisection, else ranks per element
        in FE_Element (ascending tag) order."""
        if nparts == 1:
            self.neq = self._ck(lib.xb_setup(self._h, numberer, soe))
        else:
            part = _i32(part) if part is not None else None
            self.neq = self._ck(lib.xb_setup_partitioned(self._h, numberer, soe, nparts, rank, _ptr(part)))
        self.nparts, self.rank = nparts, rank
        self.nrows = lib.xb_num_rows(self._h)
        self.nn = lib.xb_num_nodes(self._h)
        self.ne = lib.xb_num_elements(self._h)
        self.ngp = lib.xb_num_gauss_points(self._h)
        self.nnz = lib.xb_nnz(self._h)
        self.a_size = lib.xb_a_size(self._h)    # length of A: nnz, or the band / profile array
        return self.neq

    def node_tags(self):
        a = np.zeros(self.nn, np.int32); self._ck(lib.xb_get_node_tags(self._h, _ptr(a))); return a

    def ids(self):
        a = np.zeros((self.nn, self.ndf), np.int32); self._ck(lib.xb_get_ids(self._h, _ptr(a))); return a

    def element_tags(self):
        a = np.zeros(self.ne, np.int32); self._ck(lib.xb_get_element_tags(self._h, _ptr(a))); return a

    def row_eqns(self):
        a = np.zeros(self.nrows, np.int32); self._ck(lib.xb_get_row_eqns(self._h, _ptr(a))); return a

    def partition(self, ne_global):
        a = np.zeros(ne_global, np.int32); self._ck(lib.xb_get_partition(self._h, _ptr(a))); return a

    def peers(self):
        out = []
        for i in range(lib.xb_num_peers(self._h)):
            r = ctypes.c_int(0); c = np.zeros(6, np.int64)
            self._ck(lib.xb_get_peer(self._h, i, ctypes.addressof(r), _ptr(c)))
            out.append((r.value, c))
        return out

    def comm_init(self, id128: bytes):
        buf = ctypes.create_string_buffer(id128, 128)
        self._ck(lib.xb_comm_init(self._h, ctypes.addressof(buf)))

    def exchange(self, which):
        self._ck(lib.xb_exchange(self._h, which))

    def pattern(self):
        ptr = np.zeros(self.nrows + 1, np.int64); idx = np.zeros(self.nnz, np.int32)
        self._ck(lib.xb_get_pattern(self._h, _ptr(ptr), _ptr(idx))); return ptr, idx

    def scatter_map(self, e0, e1, nd):
        m = np.zeros((e1 - e0, nd, nd), np.int64)
        self._ck(lib.xb_get_scatter_map(self._h, e0, e1, _ptr(m))); return m

    # ---- device ----
    def to_device(self, device=0, stream=None):
        """stream: a raw cudaStream_t (int), e.g. torch.cuda.Stream().cuda_stream.  None (or 0, the
        legacy default stream's handle) lets the library create its own non-blocking stream; to run
        on the legacy default stream itself pass cudaStreamLegacy (1)."""
        self._ck(lib.xb_device_init(self._h, device, ctypes.c_void_p(stream) if stream else None))
        return self

    def set_trial_disp(self, u):
        u = _f64(u); assert u.size == self.nn * self.ndf
        self._ck(lib.xb_set_trial_disp(self._h, _ptr(u)))
        self._keep = u

    def incr_trial_disp(self, dU):
        dU = _f64(dU); assert dU.size == self.neq
        self._ck(lib.xb_incr_trial_disp(self._h, _ptr(dU)))
        self._keep = dU

    def trial_disp(self):
        u = np.zeros((self.nn, self.ndf)); self._ck(lib.xb_get_trial_disp(self._h, _ptr(u))); return u

    # ---- transient (Newmark) ----
    def set_mass(self, node_tags, mass):
        """`mass` command, before setup(); mass is [n][ndf] (diagonal terms)"""
        node_tags, mass = _i32(node_tags), _f64(mass)
        self._ck(lib.xb_set_nodal_mass(self._h, len(node_tags), _ptr(node_tags), _ptr(mass)))

    def set_rayleigh_alpha_m(self, alpha_m):
        self._ck(lib.xb_set_rayleigh_alpha_m(self._h, float(alpha_m)))

    def set_rayleigh(self, alpha_m, beta_k, beta_k0, beta_kc):
        """`rayleigh alphaM betaK betaKinit betaKcomm` on every element and node"""
        self._ck(lib.xb_set_rayleigh(self._h, float(alpha_m), float(beta_k), float(beta_k0), float(beta_kc)))

    def set_transient(self, c1, c2, c3):
        self._ck(lib.xb_set_transient_factors(self._h, c1, c2, c3))

    def newmark_predict(self, a1, a2, a3, a4):
        self._ck(lib.xb_newmark_predict(self._h, a1, a2, a3, a4))

    def incr_response(self, dU, cu, cv, ca, update=True):
        """Newmark::update: response increment, then (as the reference does) updateDomain"""
        dU = _f64(dU); assert dU.size == self.neq
        self._ck(lib.xb_incr_trial_response(self._h, _ptr(dU), cu, cv, ca))
        self._keep = dU
        if update:
            self.update()

    def set_vel_accel(self, v, a):
        """AnalysisModel::setVel / setAccel: trial velocities and accelerations of this rank's nodes, [nn][ndf]"""
        v, a = _f64(v), _f64(a)
        self._ck(lib.xb_set_trial_vel_accel(self._h, _ptr(v), _ptr(a)))

    def vel_accel(self):
        v = np.zeros((self.nn, self.ndf)); a = np.zeros((self.nn, self.ndf))
        self._ck(lib.xb_get_trial_vel_accel(self._h, _ptr(v), _ptr(a))); return v, a

    def update(self):
        self._ck(lib.xb_update(self._h))

    def apply_load(self, lam):
        """AnalysisModel::applyLoadDomain: the domain time / load factor and the constraint handler's applyLoad (under
        set_option("constraints_transformation", 1): the second update of the elements with a constrained node)"""
        self._ck(lib.xb_apply_load(self._h, float(lam)))

    def set_load_factor(self, lam):
        """the load factor alone (no handler action)"""
        self._ck(lib.xb_set_load_factor(self._h, float(lam)))

    def load_const(self):
        """loadConst: the loads applied so far stay at the current factor (follow with apply_load(new time))"""
        self._ck(lib.xb_load_const(self._h))

    def add_load(self, node, vals):
        self.set_nodal_loads([node], [vals])

    def set_nodal_loads(self, tags, values):
        """the next `pattern Plain`: reference loads [n][ndf] of the listed nodes"""
        tags, values = _i32(tags), _f64(values)
        self._ck(lib.xb_set_nodal_loads(self._h, len(tags), _ptr(tags), _ptr(values)))

    def form_tangent(self, out=None, host=True):
        if host and out is None:
            out = np.empty(self.a_size)
        self._ck(lib.xb_form_tangent(self._h, _ptr(out) if host else None))
        return out

    def form_unbalance(self, out=None, host=True):
        if host and out is None:
            out = np.empty(self.nrows)
        self._ck(lib.xb_form_unbalance(self._h, _ptr(out) if host else None))
        return out

    def form_element_tangents(self):
        self._ck(lib.xb_form_element_tangents(self._h))

    def assemble_tangent(self, out=None):
        self._ck(lib.xb_assemble_tangent(self._h, _ptr(out)))
        return out

    def form_element_resids(self):
        self._ck(lib.xb_form_element_resids(self._h))

    def assemble_unbalance(self, out=None):
        self._ck(lib.xb_assemble_unbalance(self._h, _ptr(out)))
        return out

    def commit(self):
        self._ck(lib.xb_commit(self._h))

    def revert_to_start(self):
        self._ck(lib.xb_revert_to_start(self._h))

    def revert_to_last_commit(self):
        self._ck(lib.xb_revert_to_last_commit(self._h))

    def synchronize(self):
        self._ck(lib.xb_synchronize(self._h))

    def element_tangent(self, e, nd):
        K = np.zeros((nd, nd)); self._ck(lib.xb_get_element_tangent(self._h, e, _ptr(K))); return K

    def element_resid(self, e, nd):
        R = np.zeros(nd); self._ck(lib.xb_get_element_resid(self._h, e, _ptr(R))); return R

    def gp_response(self, e, g, order):
        s = np.zeros(order); t = np.zeros((order, order))
        self._ck(lib.xb_get_gp_response(self._h, e, g, _ptr(s), _ptr(t))); return s, t

    def band(self):
        """BandGenLinSOE's (numSubD, numSuperD)"""
        a, b = ctypes.c_int(), ctypes.c_int()
        self._ck(lib.xb_get_band(self._h, ctypes.addressof(a), ctypes.addressof(b)))
        return a.value, b.value

    def profile(self):
        """ProfileSPDLinSOE's iDiagLoc (1-based)"""
        d = np.zeros(self.neq, np.int32)
        self._ck(lib.xb_get_profile(self._h, _ptr(d)))
        return d

    def launch_count(self):
        return lib.xb_launch_count(self._h)

    def set_option(self, name: str, value: int):
        """run-time options (include/xara_b200.h, xb_set_option): tuning switches, which do not change results, and
        "constraints_transformation", which follows the reference's `constraints Transformation`"""
        self._ck(lib.xb_set_option(self._h, name.encode(), int(value)))
        return self

    def algorithmic_bytes(self, which):
        return lib.xb_algorithmic_bytes(self._h, which)
