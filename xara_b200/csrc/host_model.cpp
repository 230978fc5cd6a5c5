// Host-side analysis set-up for the device path (see host_model.hpp for the
// reference file:line each step mirrors).
#include "host_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "../../include/xara_b200.h"

namespace xb {

static const EleKind kBrick{8, 3, 8, 6, 3};
static const EleKind kQuad{4, 2, 4, 3, 5};  // par kept: thickness, b1, b2, type (0 PlaneStrain, 1 PlaneStress), pressure
// nip is a property of the batch.  par kept per element: nIP, maxIters, tol, [vecxz[3],] then the element loads of the
// Linear pattern: `eleLoad -beamPoint` Py, Pz, N, aOverL (has-load flag in a 5th value) and `eleLoad -beamUniform` wy, wz, wa
// (zero: none) -- they travel with the element through the partitioning like every other element parameter
// (`eleLoad -beamUniform` over part of the element, Beam2d/3dPartialUniformLoad: wya, wyb, waa, wab, aOverL, bOverL, wza, wzb + has-load flag)
static const EleKind kBeam2d{2, 3, 0, 2, 25};  // par: nIP, maxIters, tol, partial load[9], joint offsets[4], rho, point load[5], wy, wz (unused), wa
static const EleKind kBeam3d{2, 6, 0, 4, 30};  // par: nIP, maxIters, tol, vecxz[3], partial load[9], joint offsets[6], rho, point load[5], wy, wz, wa

const EleKind& ele_kind(int kind) {
  return kind == XB_ELE_STDBRICK ? kBrick : (kind == XB_ELE_FOURNODEQUAD ? kQuad : (kind == XB_ELE_FORCEBEAMCOLUMN3D ? kBeam3d : kBeam2d));
}

int HostModel::add_nodes(int n, const int* tags, const double* c) {
  if (is_setup) { err = "xb_add_nodes after xb_setup"; return XB_ERR_STATE; }
  if (n < 0 || (n > 0 && (!tags || !c))) { err = "xb_add_nodes: null input"; return XB_ERR_ARG; }
  node_tag.insert(node_tag.end(), tags, tags + n);
  crd.insert(crd.end(), c, c + (size_t)n * ndm);
  return XB_OK;
}

int HostModel::add_sp(int n, const int* tags, const int* dofs) {
  if (is_setup) { err = "xb_add_sp after xb_setup"; return XB_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    if (dofs[i] < 0 || dofs[i] >= ndf) { err = "xb_add_sp: dof out of range"; return XB_ERR_ARG; }
    sp_node.push_back(tags[i]);
    sp_dof.push_back(dofs[i]);
  }
  return XB_OK;
}

int HostModel::set_node_ndf(int n, const int* tags, int nd) {
  if (is_setup) { err = "xb_set_node_ndf after xb_setup"; return XB_ERR_STATE; }
  if (nd < 1 || nd > ndf) { err = "xb_set_node_ndf: a node has between 1 and the model's ndf dofs"; return XB_ERR_ARG; }
  for (int i = 0; i < n; i++) { ndf_node.push_back(tags[i]); ndf_val.push_back(nd); }
  return XB_OK;
}

int HostModel::add_material(int tag, int kind, const double* par, int npar) {
  int need = kind == XB_MAT_J2PLASTICITY ? 7 : (kind == XB_MAT_ELASTIC_ISOTROPIC ? 2 : -1);
  if (need < 0) { err = "xb_add_nd_material: unknown kind"; return XB_ERR_ARG; }
  if (npar < need || npar > 8) { err = "xb_add_nd_material: wrong parameter count"; return XB_ERR_ARG; }
  for (auto& m : mats) if (m.tag == tag) { err = "xb_add_nd_material: duplicate tag"; return XB_ERR_ARG; }
  Material m{};
  m.tag = tag; m.kind = kind;
  std::memcpy(m.par, par, sizeof(double) * npar);
  mats.push_back(m);
  return XB_OK;
}

int HostModel::add_uniaxial(int tag, int kind, const double* par, int npar) {
  if (is_setup) { err = "xb_add_uniaxial_material after xb_setup"; return XB_ERR_STATE; }
  const int need = kind == XB_UNI_STEEL02 ? 10 : (kind == XB_UNI_CONCRETE02 ? 7 : (kind == XB_UNI_STEEL01 ? 7 : (kind == XB_UNI_ELASTIC ? 1 : ((kind == XB_UNI_CONCRETE01 || kind == XB_UNI_ELASTICPP) ? 4 : -1))));
  if (need < 0) { err = "xb_add_uniaxial_material: unknown kind"; return XB_ERR_ARG; }
  if (npar < need || npar > 12) { err = "xb_add_uniaxial_material: wrong parameter count"; return XB_ERR_ARG; }
  if (kind == XB_UNI_ELASTIC && npar > 1 && par[1] != 0.0) { err = "uniaxialMaterial Elastic with eta != 0: strain rates are outside the device path"; return XB_ERR_UNSUPPORTED; }
  for (auto& u : unis) if (u.tag == tag) { err = "xb_add_uniaxial_material: duplicate tag"; return XB_ERR_ARG; }
  Uniaxial u{};
  u.tag = tag; u.kind = kind;
  std::memcpy(u.par, par, sizeof(double) * npar);
  if (kind == XB_UNI_ELASTIC && npar < 3) u.par[2] = u.par[0];   // ElasticMaterial(tag, E, eta): Eneg = E
  if (kind == XB_UNI_ELASTICPP) {   // ElasticPPMaterial(tag, E, eyp, eyn, ezero), ElasticPPMaterial.cpp:88-107: par becomes E, fyp, fyn, ezero
    double eyp = u.par[1], eyn = u.par[2];
    if (eyp < 0) eyp *= -1.;
    if (eyn > 0) eyn *= -1.;
    u.par[1] = u.par[0] * eyp; u.par[2] = u.par[0] * eyn;
  }
  if (kind == XB_UNI_CONCRETE02 || kind == XB_UNI_CONCRETE01) {   // Concrete02.cpp:101-104, Concrete01.cpp:96-107: compression quantities are made negative
    for (int i = 0; i < 4; i++) if (u.par[i] > 0) u.par[i] = -u.par[i];
  }
  unis.push_back(u);
  return XB_OK;
}

int HostModel::add_fiber_section(int tag, int nf, const double* y, const double* A, const int* mat_tags) {
  if (is_setup) { err = "xb_add_fiber_section after xb_setup"; return XB_ERR_STATE; }
  if (nf <= 0) { err = "xb_add_fiber_section: no fibres"; return XB_ERR_ARG; }
  for (auto& sdef : secs) if (sdef.tag == tag) { err = "xb_add_fiber_section: duplicate tag"; return XB_ERR_ARG; }
  FiberSectionDef d;
  d.tag = tag;
  d.y.assign(y, y + nf); d.A.assign(A, A + nf); d.mat.resize(nf);
  double ABar = 0.0, QzBar = 0.0;
  for (int i = 0; i < nf; i++) {
    d.mat[i] = -1;
    for (size_t j = 0; j < unis.size(); j++) if (unis[j].tag == mat_tags[i]) d.mat[i] = (int)j;
    if (d.mat[i] < 0) { err = "xb_add_fiber_section: unknown uniaxial material tag"; return XB_ERR_ARG; }
    ABar += A[i]; QzBar += y[i] * A[i]; d.yBar = QzBar / ABar;     // FiberSection2d::addFiber, FiberSection2d.cpp:150-154
  }
  secs.push_back(std::move(d));
  return XB_OK;
}

int HostModel::add_section_aggregator(int tag, int n, const int* mat_tags, const int* codes) {
  if (n != 2 || codes[0] != 2 || codes[1] != 1) {
    err = "section Aggregator: the device path takes two materials with codes P, Mz in that order (a 2D forceBeamColumn section)";
    return XB_ERR_UNSUPPORTED;
  }
  const double zero[2] = {0.0, 0.0}, one[2] = {1.0, 1.0};
  int rc = add_fiber_section(tag, 2, zero, one, mat_tags);
  if (rc < 0) return rc;
  secs.back().agg = true; secs.back().yBar = 0.0;
  return XB_OK;
}

int HostModel::add_fiber_section3d(int tag, int nf, const double* y, const double* z, const double* A, const int* mat_tags, double GJ) {
  if (!z || !(GJ > 0.0)) { err = "xb_add_fiber_section3d: needs fibre z coordinates and GJ > 0 (section Fiber -GJ)"; return XB_ERR_ARG; }
  int rc = add_fiber_section(tag, nf, y, A, mat_tags);
  if (rc < 0) return rc;
  FiberSectionDef& d = secs.back();
  d.z.assign(z, z + nf); d.GJ = GJ; d.is3d = true;
  double ABar = 0.0, QyBar = 0.0;
  for (int i = 0; i < nf; i++) { ABar += A[i]; QyBar += z[i] * A[i]; d.zBar = QyBar / ABar; }   // FiberSection3d::addFiber, FiberSection3d.cpp:343-350
  return XB_OK;
}

// element dofs of one slot that land on the same column (tied dofs inside one element): the k-th of them, in
// element dof order, carries rank k in the top 3 bits of its position; the assembly adds rank 0, then 1, ...
// (the order addA meets them).  Returns false when more than 8 share a column.
static bool dup_ranks(uint16_t* cp, int n) {
  bool ok = true;
  for (int i = 0; i < n; i++) {
    if (cp[i] == 0xFFFF) continue;
    int rank = 0;
    for (int j = 0; j < i; j++) if (cp[j] != 0xFFFF && (cp[j] & 0x1FFF) == cp[i]) rank++;
    if (rank > 7) { ok = false; rank = 7; }
    cp[i] = (uint16_t)(cp[i] | (rank << 13));
  }
  return ok;
}

int HostModel::add_equal_dof(int r_tag, int c_tag, int n, const int* dofs) {
  if (is_setup) { err = "xb_add_equal_dof after xb_setup"; return XB_ERR_STATE; }
  if (r_tag == c_tag || n < 1) { err = "xb_add_equal_dof: retained and constrained node must differ, n >= 1"; return XB_ERR_ARG; }
  for (int i = 0; i < n; i++) {
    if (dofs[i] < 0 || dofs[i] >= ndf) { err = "xb_add_equal_dof: dof out of range"; return XB_ERR_ARG; }
    mp_r.push_back(r_tag); mp_c.push_back(c_tag); mp_dof.push_back(dofs[i]);
  }
  return XB_OK;
}

int HostModel::add_elements(int kind, int n, const int* tags, const int* conn, const int* mat_tags,
                            const double* par, int par_stride) {
  if (is_setup) { err = "xb_add_elements after xb_setup"; return XB_ERR_STATE; }
  if (kind != XB_ELE_STDBRICK && kind != XB_ELE_FOURNODEQUAD && kind != XB_ELE_FORCEBEAMCOLUMN2D && kind != XB_ELE_FORCEBEAMCOLUMN3D) { err = "xb_add_elements: unknown kind"; return XB_ERR_ARG; }
  const EleKind& k = ele_kind(kind);
  if (k.ndf > ndf) { err = "xb_add_elements: the element has more dofs per node than the model's ndf"; return XB_ERR_UNSUPPORTED; }
  if (kind == XB_ELE_FORCEBEAMCOLUMN2D || kind == XB_ELE_FORCEBEAMCOLUMN3D) {
    const bool b3 = kind == XB_ELE_FORCEBEAMCOLUMN3D;
    if (!b3 && (ndm != 2 || par_stride < 3)) { err = "forceBeamColumn (2D): ndm must be 2 and par = nIP, maxIters, tol"; return XB_ERR_ARG; }
    if (b3 && (ndm != 3 || par_stride < 6)) { err = "forceBeamColumn (3D): ndm must be 3 and par = nIP, maxIters, tol, vecxz[3]"; return XB_ERR_ARG; }
    Group g;
    g.kind = kind; g.mat_kind = 0;
    g.tag.assign(tags, tags + n);
    g.conn.assign(conn, conn + (size_t)n * 2);
    g.mat.assign(n, 0);
    g.par.resize((size_t)n * k.npar);
    for (int i = 0; i < n; i++) {
      const double* p = par + (size_t)i * par_stride;
      int sidx = -1;
      for (size_t j = 0; j < secs.size(); j++) if (secs[j].tag == mat_tags[i]) sidx = (int)j;
      if (sidx < 0) { err = "forceBeamColumn: unknown section tag"; return XB_ERR_ARG; }
      if (secs[sidx].is3d != b3) { err = "forceBeamColumn: a 3D element needs xb_add_fiber_section3d, a 2D one xb_add_fiber_section"; return XB_ERR_ARG; }
      // geomTransf: the parameter behind the ones above (2D: par[3], 3D: par[6]) when the caller's rows are that long
      const int tpos = b3 ? 6 : 3;
      const int transf = par_stride > tpos ? (int)p[tpos] : 0;
      if (transf != 0 && transf != 1 && !(transf == 2 && !b3)) { err = "forceBeamColumn: geomTransf is 0 (Linear), 1 (PDelta) or -- 2D -- 2 (Corotational); a 3D corotational transformation is outside the device path"; return XB_ERR_UNSUPPORTED; }
      if (i == 0) { g.sec = sidx; g.nip = (int)p[0]; g.max_iters = (int)p[1]; g.tol = p[2]; g.transf = transf; }
      else if (sidx != g.sec || (int)p[0] != g.nip || (int)p[1] != g.max_iters || p[2] != g.tol || transf != g.transf) {
        err = "forceBeamColumn: one section / nIP / maxIters / tol / geomTransf per xb_add_elements call"; return XB_ERR_UNSUPPORTED;
      }
      const int nin = b3 ? 6 : 3;            // what the caller gives; the element-load columns start at zero
      for (int q = 0; q < k.npar; q++) g.par[(size_t)i * k.npar + q] = q < nin ? p[q] : 0.0;
      // `-mass rho` (mass per unit length, lumped: ForceBeamColumn2d::getMass): the parameter behind geomTransf
      const int rpos = b3 ? 7 : 4;
      if (par_stride > rpos) g.par[(size_t)i * k.npar + k.npar - 9] = p[rpos];
      // `geomTransf ... -jntOffset dXi dYi dXj dYj` (2D): the four parameters behind rho
      if (!b3 && par_stride >= 9) for (int q = 0; q < 4; q++) g.par[(size_t)i * k.npar + k.npar - 13 + q] = p[5 + q];
      // 3D: -jntOffset dXi dYi dZi dXj dYj dZj, parameters 8..13
      if (b3 && par_stride >= 14) for (int q = 0; q < 6; q++) g.par[(size_t)i * k.npar + k.npar - 15 + q] = p[8 + q];
    }
    if (g.nip < 2 || g.nip > 10) { err = "forceBeamColumn: Lobatto integration takes 2..10 points"; return XB_ERR_ARG; }
    if (n > 0) groups.push_back(std::move(g));
    return XB_OK;
  }
  if ((kind == XB_ELE_STDBRICK && (ndm != 3 || par_stride < 3)) ||
      (kind == XB_ELE_FOURNODEQUAD && (ndm != 2 || par_stride < 6))) {
    err = "xb_add_elements: ndm / par_stride do not fit the element kind"; return XB_ERR_ARG;
  }
  Group g;
  g.kind = kind;
  g.tag.assign(tags, tags + n);
  g.conn.assign(conn, conn + (size_t)n * k.nen);
  g.mat.resize(n);
  g.par.resize((size_t)n * k.npar);
  int mk = -1;
  // material tag -> index (few materials: linear probe with a one-entry cache)
  int last_tag = 0, last_idx = -1;
  for (int i = 0; i < n; i++) {
    int mt = mat_tags[i];
    if (last_idx < 0 || mt != last_tag) {
      last_idx = -1;
      for (size_t j = 0; j < mats.size(); j++) if (mats[j].tag == mt) { last_idx = (int)j; break; }
      if (last_idx < 0) { err = "xb_add_elements: unknown material tag"; return XB_ERR_ARG; }
      last_tag = mt;
    }
    g.mat[i] = last_idx;
    if (mk < 0) mk = mats[last_idx].kind;
    else if (mk != mats[last_idx].kind) { err = "xb_add_elements: one nDMaterial kind per call"; return XB_ERR_UNSUPPORTED; }
    const double* p = par + (size_t)i * par_stride;
    double* q = &g.par[(size_t)i * k.npar];
    if (kind == XB_ELE_STDBRICK) { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; }
    else {
      if ((int)p[1] != 0 && (int)p[1] != 1) { err = "FourNodeQuad: type is 0 (PlaneStrain) or 1 (PlaneStress)"; return XB_ERR_ARG; }
      // J2Plasticity's PlaneStress copy is another class (J2PlaneStress) with a state of its own (the out-of-plane strain):
      // one plane type per batch of J2 quads, so that the batch's commit / revert can treat that state as a block
      if (mats[last_idx].kind == XB_MAT_J2PLASTICITY) {
        if (i == 0) g.j2_plane_stress = (int)p[1] == 1;
        else if (g.j2_plane_stress != ((int)p[1] == 1)) { err = "FourNodeQuad with J2Plasticity: one plane type (PlaneStrain | PlaneStress) per xb_add_elements call"; return XB_ERR_UNSUPPORTED; }
      }
      // par[3] is the element's own density, which FourNodeQuad uses instead of the material's when non-zero
      // (FourNodeQuad.cpp:395-398): the device kernels take the material density only
      if (p[3] != 0.0) { err = "FourNodeQuad: an element density (rho) is not supported; give the density on the nDMaterial"; return XB_ERR_UNSUPPORTED; }
      q[0] = p[0]; q[1] = p[4]; q[2] = p[5]; q[3] = (double)(int)p[1]; q[4] = p[2];
    }
  }
  g.mat_kind = mk < 0 ? 0 : mk;
  if (n > 0) groups.push_back(std::move(g));
  return XB_OK;
}

// `eleLoad -ele tags -type -beamUniform wy [wz] wa` (Beam2dUniformLoad / Beam3dUniformLoad) in the Linear pattern:
// ForceBeamColumn2d/3d::addLoad.  w: [n][3] = wy, wz, wa.  One uniform load per element.
int HostModel::add_beam_uniform_loads(int n, const int* tags, const double* w) {
  if (is_setup) { err = "xb_add_beam_uniform_loads after xb_setup"; return XB_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    bool found = false;
    for (auto& g : groups) {
      if (g.kind != XB_ELE_FORCEBEAMCOLUMN2D && g.kind != XB_ELE_FORCEBEAMCOLUMN3D) continue;
      const int npar = ele_kind(g.kind).npar;
      for (size_t l = 0; l < g.tag.size() && !found; l++) {
        if (g.tag[l] != tags[i]) continue;
        double* q = &g.par[l * npar + npar - 3];
        // several uniform loads on one element (dead + live): the element adds their section forces and reactions, all
        // at the pattern's load factor (ForceBeamColumn2d.cpp:1034-1070) -- one load with the summed intensities
        q[0] += w[(size_t)i * 3]; q[1] += g.kind == XB_ELE_FORCEBEAMCOLUMN3D ? w[(size_t)i * 3 + 1] : 0.0; q[2] += w[(size_t)i * 3 + 2];
        found = true;
      }
      if (found) break;
    }
    if (!found) { err = "xb_add_beam_uniform_loads: no forceBeamColumn element with this tag"; return XB_ERR_ARG; }
  }
  return XB_OK;
}

int HostModel::add_loads(int n, const int* tags, const double* vals) {
  if (is_setup) { err = "xb_add_nodal_loads after xb_setup"; return XB_ERR_STATE; }
  load_node.insert(load_node.end(), tags, tags + n);
  load_val.insert(load_val.end(), vals, vals + (size_t)n * ndf);
  return XB_OK;
}

int HostModel::set_beam_integration(int n, const int* tags, int nip, const double* xi, const double* wt) {
  if (is_setup) { err = "xb_set_beam_integration after xb_setup"; return XB_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    bool found = false;
    for (auto& g : groups) {
      if (g.kind != XB_ELE_FORCEBEAMCOLUMN2D && g.kind != XB_ELE_FORCEBEAMCOLUMN3D) continue;
      for (size_t l = 0; l < g.tag.size() && !found; l++) {
        if (g.tag[l] != tags[i]) continue;
        found = true;
        if (nip != g.nip) { err = "xb_set_beam_integration: nip differs from the element's"; return XB_ERR_ARG; }
        if (g.rule.empty()) { g.rule.assign(g.tag.size() * 2 * (size_t)nip, 0.0); g.rule_set.assign(g.tag.size(), 0); }
        std::memcpy(&g.rule[l * 2 * nip], xi + (size_t)i * nip, sizeof(double) * nip);
        std::memcpy(&g.rule[l * 2 * nip + nip], wt + (size_t)i * nip, sizeof(double) * nip);
        g.rule_set[l] = 1;
      }
      if (found) break;
    }
    if (!found) { err = "xb_set_beam_integration: no forceBeamColumn element with this tag"; return XB_ERR_ARG; }
  }
  return XB_OK;
}

int HostModel::add_beam_partial_loads(int n, const int* tags, const double* pv) {
  if (is_setup) { err = "xb_add_beam_partial_loads after xb_setup"; return XB_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    const double* q8 = pv + (size_t)i * 8;
    if (!(q8[4] >= 0.0 && q8[4] < q8[5] && q8[5] <= 1.0)) { err = "xb_add_beam_partial_loads: 0 <= aOverL < bOverL <= 1"; return XB_ERR_ARG; }
    bool found = false;
    for (auto& g : groups) {
      if (g.kind != XB_ELE_FORCEBEAMCOLUMN2D && g.kind != XB_ELE_FORCEBEAMCOLUMN3D) continue;
      const bool b3 = g.kind == XB_ELE_FORCEBEAMCOLUMN3D;
      const int npar = ele_kind(g.kind).npar;
      for (size_t l = 0; l < g.tag.size() && !found; l++) {
        if (g.tag[l] != tags[i]) continue;
        found = true;
        double* q = &g.par[l * npar + (b3 ? 6 : 3)];
        if (q[8] != 0.0) { err = "xb_add_beam_partial_loads: one partial uniform load per element"; return XB_ERR_UNSUPPORTED; }
        for (int c = 0; c < 8; c++) q[c] = (c < 6 || b3) ? q8[c] : 0.0;
        q[8] = 1.0;
      }
      if (found) break;
    }
    if (!found) { err = "xb_add_beam_partial_loads: no forceBeamColumn element with this tag"; return XB_ERR_ARG; }
  }
  return XB_OK;
}

int HostModel::add_beam_point_loads(int n, const int* tags, const double* pv) {
  if (is_setup) { err = "xb_add_beam_point_loads after xb_setup"; return XB_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    const double* q4 = pv + (size_t)i * 4;
    bool found = false;
    for (auto& g : groups) {
      if (g.kind != XB_ELE_FORCEBEAMCOLUMN2D && g.kind != XB_ELE_FORCEBEAMCOLUMN3D) continue;
      const int npar = ele_kind(g.kind).npar;
      for (size_t l = 0; l < g.tag.size() && !found; l++) {
        if (g.tag[l] != tags[i]) continue;
        found = true;
        if (q4[3] < 0.0 || q4[3] > 1.0) break;      // the element ignores such a load (ForceBeamColumn2d.cpp:447)
        double* q = &g.par[l * npar + npar - 8];
        if (q[4] != 0.0) { err = "xb_add_beam_point_loads: one point load per element"; return XB_ERR_UNSUPPORTED; }
        q[0] = q4[0]; q[1] = g.kind == XB_ELE_FORCEBEAMCOLUMN3D ? q4[1] : 0.0; q[2] = q4[2]; q[3] = q4[3]; q[4] = 1.0;
      }
      if (found) break;
    }
    if (!found) { err = "xb_add_beam_point_loads: no forceBeamColumn element with this tag"; return XB_ERR_ARG; }
  }
  return XB_OK;
}

int HostModel::add_mass(int n, const int* tags, const double* vals) {
  if (is_setup) { err = "xb_set_nodal_mass after xb_setup"; return XB_ERR_STATE; }
  mass_node.insert(mass_node.end(), tags, tags + n);
  mass_val.insert(mass_val.end(), vals, vals + (size_t)n * ndf);
  return XB_OK;
}

namespace {
struct TagIndex {
  const std::vector<int>& tags;
  bool contiguous;
  int base;
  explicit TagIndex(const std::vector<int>& t) : tags(t), contiguous(true), base(t.empty() ? 0 : t[0]) {
    for (size_t i = 0; i < t.size(); i++) if (t[i] != base + (int)i) { contiguous = false; break; }
  }
  int operator()(int tag) const {
    if (contiguous) { long long i = (long long)tag - base; return (i >= 0 && i < (long long)tags.size()) ? (int)i : -1; }
    auto it = std::lower_bound(tags.begin(), tags.end(), tag);
    return (it != tags.end() && *it == tag) ? (int)(it - tags.begin()) : -1;
  }
};
}  // namespace

namespace {
// global connectivity view used while setting up
struct GlobalMesh {
  const std::vector<Group>* groups;
  std::vector<int> fe_group, fe_local;       // [ne] FE order
  std::vector<long long> n2e_ptr;            // [nn+1]
  std::vector<int> n2e_fe;                   // [*] FE index
  std::vector<uint8_t> n2e_loc;              // [*]
  const int* conn_of(long long e, const EleKind** k) const {
    const Group& g = (*groups)[fe_group[e]];
    *k = &ele_kind(g.kind);
    return &g.conn[(size_t)fe_local[e] * (*k)->nen];
  }
  // neighbours of node n (sorted, unique, including n when it has an element)
  void nbrs(int n, std::vector<int>& out) const {
    out.clear();
    for (long long t = n2e_ptr[n]; t < n2e_ptr[n + 1]; t++) {
      const EleKind* k;
      const int* c = conn_of(n2e_fe[t], &k);
      out.insert(out.end(), c, c + k->nen);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
  }
};

// recursive coordinate bisection of element centroids into `np` parts, deterministic
void rcb(std::vector<long long>& items, long long lo, long long hi, int np, int first, const double* cen,
         int ndm, std::vector<int>& part) {
  if (np == 1) { for (long long i = lo; i < hi; i++) part[items[i]] = first; return; }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (long long i = lo; i < hi; i++)
    for (int d = 0; d < ndm; d++) { double v = cen[items[i] * 3 + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
  int ax = 0;
  for (int d = 1; d < ndm; d++) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  const int nl = np / 2;
  const long long k = lo + (hi - lo) * nl / np;
  std::nth_element(items.begin() + lo, items.begin() + k, items.begin() + hi, [&](long long a, long long b) {
    double va = cen[a * 3 + ax], vb = cen[b * 3 + ax];
    return va < vb || (va == vb && a < b);
  });
  rcb(items, lo, k, nl, first, cen, ndm, part);
  rcb(items, k, hi, np - nl, first + nl, cen, ndm, part);
}
}  // namespace

int HostModel::setup(int numberer_, int soe_kind_, int nparts_, int rank_, const int* part_in) {
  for (const Group& g : groups)
    for (uint8_t f : g.rule_set) if (!f) { err = "xb_set_beam_integration: every element of an xb_add_elements call, or none"; return XB_ERR_ARG; }
  if (is_setup) { err = "xb_setup called twice"; return XB_ERR_STATE; }
  if (numberer_ != XB_NUMBERER_PLAIN && numberer_ != XB_NUMBERER_RCM) { err = "unknown numberer"; return XB_ERR_ARG; }
  if (soe_kind_ < XB_SOE_SPARSE_GEN_COL || soe_kind_ > XB_SOE_UMFPACK_GEN) { err = "unknown SOE kind"; return XB_ERR_ARG; }
  soe_store = soe_kind_;
  if (soe_kind_ != XB_SOE_SPARSE_GEN_ROW) soe_kind_ = XB_SOE_SPARSE_GEN_COL;   // band / profile / Umfpack: column-oriented addA
  if (nparts_ > 1 && (soe_store == XB_SOE_BAND_GEN || soe_store == XB_SOE_PROFILE_SPD)) {
    err = "BandGeneral / ProfileSPD storage is for single-GPU models"; return XB_ERR_UNSUPPORTED;
  }
  if (nparts_ < 1 || rank_ < 0 || rank_ >= nparts_) { err = "bad nparts / rank"; return XB_ERR_ARG; }
  numberer = numberer_; soe_kind = soe_kind_; nparts = nparts_; rank = rank_;
  const int n_nodes = (int)node_tag.size();
  nn_global = n_nodes;

  // ---- Domain node map iterates by ascending tag (MapOfTaggedObjects, Domain.cpp:98) ----
  if (!std::is_sorted(node_tag.begin(), node_tag.end())) {
    std::vector<int> perm(n_nodes);
    std::iota(perm.begin(), perm.end(), 0);
    std::sort(perm.begin(), perm.end(), [&](int a, int b) { return node_tag[a] < node_tag[b]; });
    std::vector<int> t(n_nodes); std::vector<double> c((size_t)n_nodes * ndm);
    for (int i = 0; i < n_nodes; i++) {
      t[i] = node_tag[perm[i]];
      std::memcpy(&c[(size_t)i * ndm], &crd[(size_t)perm[i] * ndm], sizeof(double) * ndm);
    }
    node_tag.swap(t); crd.swap(c);
  }
  for (int i = 1; i < n_nodes; i++) if (node_tag[i] == node_tag[i - 1]) { err = "duplicate node tag"; return XB_ERR_ARG; }
  TagIndex nidx(node_tag);

  // ---- element connectivity: tags -> (global) node indices ----
  cp_stride = 0;
  long long bad = 0;
  for (auto& g : groups) {
    const EleKind& k = ele_kind(g.kind);
    cp_stride = std::max(cp_stride, k.nen * k.ndf);   // <= 32: one lane per element dof in assemble_A
    const long long m = (long long)g.conn.size();
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long i = 0; i < m; i++) {
      int ix = nidx(g.conn[i]);
      if (ix < 0) bad++;
      g.conn[i] = ix;
    }
  }
  if (bad) { err = "element references an unknown node tag"; return XB_ERR_ARG; }

  // ---- PlainHandler::handle: every dof -2 (free) unless an SP_Constraint sets -1 ----
  std::vector<int> gid((size_t)n_nodes * ndf, -2);
  for (size_t i = 0; i < sp_node.size(); i++) {
    int ix = nidx(sp_node[i]);
    if (ix < 0) { err = "fix references an unknown node tag"; return XB_ERR_ARG; }
    gid[(size_t)ix * ndf + sp_dof[i]] = -1;
  }
  // nodes with fewer dofs than the model's ndf: the missing ones never exist (no equation)
  std::vector<uint8_t> nndf(n_nodes, (uint8_t)ndf);
  for (size_t i = 0; i < ndf_node.size(); i++) {
    const int ix = nidx(ndf_node[i]);
    if (ix < 0) { err = "xb_set_node_ndf references an unknown node tag"; return XB_ERR_ARG; }
    nndf[ix] = (uint8_t)ndf_val[i];
    for (int j = ndf_val[i]; j < ndf; j++) gid[(size_t)ix * ndf + j] = -1;
  }
  // an element's nodes carry exactly the element's dofs per node (FourNodeQuad.cpp:133-139, Brick, ForceBeamColumn2d/3d)
  for (auto& g : groups) {
    const EleKind& k = ele_kind(g.kind);
    long long wrong = 0;
    const long long mconn = (long long)g.conn.size();
#pragma omp parallel for reduction(+ : wrong) schedule(static)
    for (long long i = 0; i < mconn; i++) if (nndf[g.conn[i]] != k.ndf) wrong++;
    if (wrong) { err = "an element is connected to a node whose number of dofs differs from the element's (xb_set_node_ndf)"; return XB_ERR_ARG; }
  }
  for (size_t i = 0; i < sp_node.size(); i++) {
    const int ix = nidx(sp_node[i]);
    if (ix >= 0 && sp_dof[i] >= nndf[ix]) { err = "fix: dof beyond the node's number of dofs"; return XB_ERR_ARG; }
  }
  // MP_Constraints with an identity matrix (equalDOF): constrained dofs get -4 unless already constrained
  // (PlainHandler.cpp:129-176 warns and keeps the SP)
  const bool have_mp = !mp_r.empty();
  std::vector<int> mp_ri(mp_r.size()), mp_ci(mp_r.size());
  for (size_t i = 0; i < mp_r.size(); i++) {
    mp_ri[i] = nidx(mp_r[i]); mp_ci[i] = nidx(mp_c[i]);
    if (mp_ri[i] < 0 || mp_ci[i] < 0) { err = "equalDOF references an unknown node tag"; return XB_ERR_ARG; }
  }
  for (size_t i = 0; i < mp_r.size(); i++) {
    int& idc = gid[(size_t)mp_ci[i] * ndf + mp_dof[i]];
    if (idc == -2) idc = -4;
  }
  for (size_t i = 0; i < mp_r.size(); i++)
    if (gid[(size_t)mp_ri[i] * ndf + mp_dof[i]] == -4) { err = "equalDOF: a retained dof is itself constrained by another equalDOF"; return XB_ERR_UNSUPPORTED; }
  std::vector<double> gload((size_t)n_nodes * ndf, 0.0);
  for (size_t i = 0; i < load_node.size(); i++) {
    int ix = nidx(load_node[i]);
    if (ix < 0) { err = "load references an unknown node tag"; return XB_ERR_ARG; }
    for (int j = 0; j < ndf; j++) gload[(size_t)ix * ndf + j] += load_val[i * ndf + j];
  }

  std::vector<double> gmass((size_t)n_nodes * ndf, 0.0);
  for (size_t i = 0; i < mass_node.size(); i++) {
    int ix = nidx(mass_node[i]);
    if (ix < 0) { err = "mass references an unknown node tag"; return XB_ERR_ARG; }
    for (int j = 0; j < ndf; j++) gmass[(size_t)ix * ndf + j] = mass_val[i * ndf + j];   // Node::setMass replaces
  }
  // forceBeamColumn -mass rho: the element's mass matrix is lumped (ForceBeamColumn2d::getMass, 3d: 0.5 rho L on the
  // translational dofs of both nodes; the rayleigh command gives elements and nodes one alphaM), i.e. nodal masses
  for (auto& g : groups) {
    if (g.kind != XB_ELE_FORCEBEAMCOLUMN2D && g.kind != XB_ELE_FORCEBEAMCOLUMN3D) continue;
    const int npar = ele_kind(g.kind).npar;
    for (long long l = 0; l < g.n(); l++) {
      const double rho = g.par[(size_t)l * npar + npar - 9];
      if (rho == 0.0) continue;
      const int a = g.conn[(size_t)l * 2], b = g.conn[(size_t)l * 2 + 1];
      double L2 = 0.0;
      for (int d = 0; d < ndm; d++) {
        double dx = crd[(size_t)b * ndm + d] - crd[(size_t)a * ndm + d];
        // (between the offset ends)
        if (g.kind == XB_ELE_FORCEBEAMCOLUMN2D) dx += g.par[(size_t)l * npar + npar - 11 + d] - g.par[(size_t)l * npar + npar - 13 + d];
        else dx += g.par[(size_t)l * npar + npar - 12 + d] - g.par[(size_t)l * npar + npar - 15 + d];
        L2 += dx * dx;
      }
      const double mL = 0.5 * rho * std::sqrt(L2);
      for (int d = 0; d < ndm; d++) { gmass[(size_t)a * ndf + d] += mL; gmass[(size_t)b * ndf + d] += mL; }
    }
  }

  // ---- FE_Element order: Domain element map by ascending tag (PlainHandler.cpp:228-250) ----
  GlobalMesh G;
  G.groups = &groups;
  long long neg = 0;
  for (auto& g : groups) neg += g.n();
  ne_global = neg;
  G.fe_group.resize(neg); G.fe_local.resize(neg);
  {
    bool simple = groups.size() == 1 && std::is_sorted(groups[0].tag.begin(), groups[0].tag.end());
    if (simple) {
      for (long long e = 0; e < neg; e++) { G.fe_group[e] = 0; G.fe_local[e] = (int)e; }
      for (size_t i = 1; i < groups[0].tag.size(); i++)
        if (groups[0].tag[i] == groups[0].tag[i - 1]) { err = "duplicate element tag"; return XB_ERR_ARG; }
    } else {
      std::vector<long long> order(neg);
      std::vector<int> tg(neg), gg(neg), ll(neg);
      long long c = 0;
      for (size_t gi = 0; gi < groups.size(); gi++)
        for (long long l = 0; l < groups[gi].n(); l++) { tg[c] = groups[gi].tag[l]; gg[c] = (int)gi; ll[c] = (int)l; c++; }
      std::iota(order.begin(), order.end(), 0LL);
      std::sort(order.begin(), order.end(), [&](long long a, long long b) { return tg[a] < tg[b]; });
      for (long long e = 0; e < neg; e++) { G.fe_group[e] = gg[order[e]]; G.fe_local[e] = ll[order[e]]; }
      for (long long e = 1; e < neg; e++) if (tg[order[e]] == tg[order[e - 1]]) { err = "duplicate element tag"; return XB_ERR_ARG; }
    }
  }

  // ---- node -> FE elements (global, FE order) ----
  G.n2e_ptr.assign((size_t)n_nodes + 1, 0);
  for (long long e = 0; e < neg; e++) {
    const EleKind* k; const int* c = G.conn_of(e, &k);
    for (int a = 0; a < k->nen; a++) G.n2e_ptr[c[a] + 1]++;
  }
  for (int n = 0; n < n_nodes; n++) G.n2e_ptr[n + 1] += G.n2e_ptr[n];
  G.n2e_fe.resize(G.n2e_ptr[n_nodes]); G.n2e_loc.resize(G.n2e_ptr[n_nodes]);
  {
    std::vector<long long> fill(G.n2e_ptr.begin(), G.n2e_ptr.end() - 1);
    for (long long e = 0; e < neg; e++) {
      const EleKind* k; const int* c = G.conn_of(e, &k);
      for (int a = 0; a < k->nen; a++) { long long t = fill[c[a]]++; G.n2e_fe[t] = (int)e; G.n2e_loc[t] = (uint8_t)a; }
    }
  }

  // ---- numbering (global, identical on every rank) ----
  std::vector<int> order(n_nodes);
  if (numberer == XB_NUMBERER_PLAIN) {
    std::iota(order.begin(), order.end(), 0);
  } else {
    // DOF_Group graph (AnalysisModel.cpp:355-400): vertex per DOF_Group, tag = position in
    // node-tag order; adjacency kept sorted by ID::insert.  RCM::number with GPS off and
    // no start vertex (RCM.cpp:186-262): BFS from the first vertex, filling the result
    // from the back; disconnected pieces restart at the next unvisited vertex.
    std::vector<long long> nb_ptr((size_t)n_nodes + 1, 0);
#pragma omp parallel
    {
      std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
      for (int n = 0; n < n_nodes; n++) { G.nbrs(n, tmp); nb_ptr[n + 1] = (long long)tmp.size(); }
    }
    for (int n = 0; n < n_nodes; n++) nb_ptr[n + 1] += nb_ptr[n];
    std::vector<int> nb(nb_ptr[n_nodes]);
#pragma omp parallel
    {
      std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
      for (int n = 0; n < n_nodes; n++) { G.nbrs(n, tmp); std::copy(tmp.begin(), tmp.end(), nb.begin() + nb_ptr[n]); }
    }
    std::vector<int> mark(n_nodes, -1);
    if (n_nodes > 0) {
      int currentMark = n_nodes - 1, nextMark = currentMark - 1, iter = 0;
      order[currentMark] = 0; mark[0] = currentMark;
      while (nextMark >= 0) {
        int v = order[currentMark];
        for (long long a = nb_ptr[v]; a < nb_ptr[v + 1]; a++) {
          int w = nb[a];
          if (w != v && mark[w] == -1) { mark[w] = nextMark; order[nextMark--] = w; }
        }
        currentMark--;
        if (currentMark == nextMark && currentMark >= 0) {
          while (iter < n_nodes && mark[iter] != -1) iter++;
          nextMark--;
          mark[iter] = currentMark; order[currentMark] = iter; iter++;
        }
      }
    }
  }
  int eqn = 0;
  for (int i = 0; i < n_nodes; i++) {
    int n = order[i];
    for (int j = 0; j < ndf; j++) if (gid[(size_t)n * ndf + j] == -2) gid[(size_t)n * ndf + j] = eqn++;
  }
  neq = eqn;
  { std::vector<int>().swap(order); }
  // the numberer's last pass (PlainNumberer.cpp:111-142, DOF_Numberer.cpp:151-190): -4 -> the retained dof's id
  std::vector<uint8_t> gshared(have_mp ? (size_t)n_nodes * ndf : 0, 0);   // (node, dof) on an equation several dofs share
  for (size_t i = 0; i < mp_r.size(); i++) {
    int& idc = gid[(size_t)mp_ci[i] * ndf + mp_dof[i]];
    if (idc != -4) continue;
    idc = gid[(size_t)mp_ri[i] * ndf + mp_dof[i]];
    if (idc >= 0) { gshared[(size_t)mp_ci[i] * ndf + mp_dof[i]] = 1; gshared[(size_t)mp_ri[i] * ndf + mp_dof[i]] = 1; }
  }

  // ---- element partition and node ownership (DomainPartitioner's role; the numbering above
  // is untouched, so every rank's rows are rows of the one global system) ----
  part_fe.assign(neg, 0);
  if (nparts > 1) {
    if (part_in) {
      for (long long e = 0; e < neg; e++) {
        if (part_in[e] < 0 || part_in[e] >= nparts) { err = "partition entry out of range"; return XB_ERR_ARG; }
        part_fe[e] = part_in[e];
      }
    } else {
      std::vector<double> cen((size_t)neg * 3, 0.0);
#pragma omp parallel for schedule(static)
      for (long long e = 0; e < neg; e++) {
        const EleKind* k; const int* c = G.conn_of(e, &k);
        for (int a = 0; a < k->nen; a++)
          for (int d = 0; d < ndm; d++) cen[e * 3 + d] += crd[(size_t)c[a] * ndm + d];
        for (int d = 0; d < ndm; d++) cen[e * 3 + d] /= k->nen;
      }
      std::vector<long long> items(neg);
      std::iota(items.begin(), items.end(), 0LL);
      rcb(items, 0, neg, nparts, 0, cen.data(), ndm, part_fe);
    }
  }
  // a node's equations belong to the lowest rank holding one of its elements
  std::vector<int> owner(n_nodes, 0);
  if (nparts > 1) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < n_nodes; n++) {
      int o = nparts;
      for (long long t = G.n2e_ptr[n]; t < G.n2e_ptr[n + 1]; t++) o = std::min(o, part_fe[G.n2e_fe[t]]);
      owner[n] = (o == nparts) ? 0 : o;
    }
    // nodes tied by an equalDOF share equations: they go to ONE rank (the lowest owner of the tie group), which then
    // holds every slot that feeds a shared row -- local ones, and remote ones through the ordinary exchange
    if (have_mp) {
      std::vector<int> root(n_nodes);
      std::iota(root.begin(), root.end(), 0);
      auto find = [&](int x) { while (root[x] != x) { root[x] = root[root[x]]; x = root[x]; } return x; };
      for (size_t i = 0; i < mp_ri.size(); i++) {
        const int a = find(mp_ri[i]), b = find(mp_ci[i]);
        if (a != b) root[std::max(a, b)] = std::min(a, b);
      }
      std::vector<int> gown(n_nodes, nparts);
      for (size_t i = 0; i < mp_ri.size(); i++)
        for (int x : {mp_ri[i], mp_ci[i]}) { const int r = find(x); gown[r] = std::min(gown[r], owner[x]); }
      for (size_t i = 0; i < mp_ri.size(); i++)
        for (int x : {mp_ri[i], mp_ci[i]}) owner[x] = gown[find(x)];
    }
  }

  // ---- local nodes: nodes of local elements and owned nodes, ascending global index ----
  std::vector<int> g2l(n_nodes, -1);
  std::vector<int> lnode;
  if (nparts == 1) {
    lnode.resize(n_nodes);
    std::iota(lnode.begin(), lnode.end(), 0);
    std::iota(g2l.begin(), g2l.end(), 0);
  } else {
    std::vector<uint8_t> mark(n_nodes, 0);
    for (int n = 0; n < n_nodes; n++) if (owner[n] == rank) mark[n] = 1;
    for (long long e = 0; e < neg; e++) if (part_fe[e] == rank) {
      const EleKind* k; const int* c = G.conn_of(e, &k);
      for (int a = 0; a < k->nen; a++) mark[c[a]] = 1;
    }
    for (int n = 0; n < n_nodes; n++) if (mark[n]) { g2l[n] = (int)lnode.size(); lnode.push_back(n); }
  }
  const int nl = (int)lnode.size();

  // ---- local element groups (subset of every batch, FE order preserved) ----
  std::vector<Group> lgroups;
  std::vector<long long> g_fe_to_local(nparts == 1 ? 0 : neg, -1);   // global FE -> local FE
  fe_group.clear(); fe_local.clear(); fe_global.clear();
  if (nparts == 1) {
    fe_group = G.fe_group; fe_local = G.fe_local;
    fe_global.resize(neg); std::iota(fe_global.begin(), fe_global.end(), 0LL);
    ne = neg;
  } else {
    lgroups.resize(groups.size());
    std::vector<std::vector<int>> keep(groups.size());   // batch index -> kept element indices
    for (size_t gi = 0; gi < groups.size(); gi++) {
      lgroups[gi].kind = groups[gi].kind; lgroups[gi].mat_kind = groups[gi].mat_kind; lgroups[gi].j2_plane_stress = groups[gi].j2_plane_stress;
      lgroups[gi].sec = groups[gi].sec; lgroups[gi].nip = groups[gi].nip; lgroups[gi].max_iters = groups[gi].max_iters; lgroups[gi].tol = groups[gi].tol;
      lgroups[gi].transf = groups[gi].transf;
    }
    std::vector<std::vector<int>> newidx(groups.size());
    for (size_t gi = 0; gi < groups.size(); gi++) newidx[gi].assign(groups[gi].n(), -1);
    // kept elements of a batch keep their batch order
    std::vector<std::vector<uint8_t>> mine(groups.size());
    for (size_t gi = 0; gi < groups.size(); gi++) mine[gi].assign(groups[gi].n(), 0);
    for (long long e = 0; e < neg; e++) if (part_fe[e] == rank) mine[G.fe_group[e]][G.fe_local[e]] = 1;
    for (size_t gi = 0; gi < groups.size(); gi++) {
      const Group& g = groups[gi]; Group& lg = lgroups[gi];
      const EleKind& k = ele_kind(g.kind);
      for (long long l = 0; l < g.n(); l++) if (mine[gi][l]) {
        newidx[gi][l] = (int)lg.tag.size();
        lg.tag.push_back(g.tag[l]); lg.mat.push_back(g.mat[l]);
        for (int a = 0; a < k.nen; a++) lg.conn.push_back(g2l[g.conn[(size_t)l * k.nen + a]]);
        for (int q = 0; q < k.npar; q++) lg.par.push_back(g.par[(size_t)l * k.npar + q]);
        if (!g.rule.empty()) { lg.rule.insert(lg.rule.end(), &g.rule[(size_t)l * 2 * g.nip], &g.rule[(size_t)(l + 1) * 2 * g.nip]); lg.rule_set.push_back(g.rule_set[l]); }
      }
    }
    ne = 0;
    for (long long e = 0; e < neg; e++) if (part_fe[e] == rank) {
      g_fe_to_local[e] = ne++;
      fe_group.push_back(G.fe_group[e]); fe_local.push_back(newidx[G.fe_group[e]][G.fe_local[e]]); fe_global.push_back(e);
    }
  }
  const std::vector<Group>& LG = nparts == 1 ? groups : lgroups;
  // offsets of the local element matrices / residuals
  std::vector<long long> ke_off(LG.size()), re_off(LG.size()), gp_off(LG.size()), rec_off(LG.size(), 0);
  ke_total = re_total = ngp = rec_total = 0;
  // stdBrick tangents are kept as symmetric element records and gathered by the assembly (brick_rec.hpp).  A model
  // with ndf = 3 in 3D holds no other element kind here, so this is the only brick path.
  rec_mode = brick_records && !groups.empty() && ndf == 3 && cp_stride == 24;     // (the global batches: the same answer on every rank)
  for (const Group& g : groups) if (g.kind != XB_ELE_STDBRICK) rec_mode = false;
  for (size_t gi = 0; gi < LG.size(); gi++) {
    const EleKind& k = ele_kind(LG[gi].kind);
    const long long nd = k.nen * k.ndf;
    ke_off[gi] = ke_total; re_off[gi] = re_total; gp_off[gi] = ngp; rec_off[gi] = rec_total;
    ke_total += LG[gi].n() * nd * nd; re_total += LG[gi].n() * nd; ngp += LG[gi].n() * (k.nip ? k.nip : LG[gi].nip);
    if (rec_mode) rec_total += LG[gi].n() * kBrickRec;
  }

  // ---- local node tables ----
  id.resize((size_t)nl * ndf); load.resize((size_t)nl * ndf); owned.assign(nl, 0); mass.resize((size_t)nl * ndf);
  {
    std::vector<int> t(nl); std::vector<double> c((size_t)nl * ndm);
    for (int i = 0; i < nl; i++) {
      const int n = lnode[i];
      t[i] = node_tag[n];
      std::memcpy(&c[(size_t)i * ndm], &crd[(size_t)n * ndm], sizeof(double) * ndm);
      for (int j = 0; j < ndf; j++) {
        id[(size_t)i * ndf + j] = gid[(size_t)n * ndf + j]; load[(size_t)i * ndf + j] = gload[(size_t)n * ndf + j];
        mass[(size_t)i * ndf + j] = gmass[(size_t)n * ndf + j];
      }
      owned[i] = owner[n] == rank;
    }
    node_tag.swap(t); crd.swap(c);
  }
  // owned rows, ascending global equation number
  row_geq.clear();
  for (int i = 0; i < nl; i++) if (owned[i])
    for (int j = 0; j < ndf; j++) if (id[(size_t)i * ndf + j] >= 0) row_geq.push_back(id[(size_t)i * ndf + j]);
  std::sort(row_geq.begin(), row_geq.end());
  if (have_mp) row_geq.erase(std::unique(row_geq.begin(), row_geq.end()), row_geq.end());
  nrows = (int)row_geq.size();
  row_of.assign((size_t)nl * ndf, -1);
  for (int i = 0; i < nl; i++) if (owned[i])
    for (int j = 0; j < ndf; j++) {
      const int q = id[(size_t)i * ndf + j];
      if (q >= 0) row_of[(size_t)i * ndf + j] = (int)(std::lower_bound(row_geq.begin(), row_geq.end(), q) - row_geq.begin());
    }

  // ---- slots of the owned nodes: every adjacent element (local or remote) in global FE order ----
  n2e_ptr.assign((size_t)nl + 1, 0);
  for (int i = 0; i < nl; i++) if (owned[i]) n2e_ptr[i + 1] = G.n2e_ptr[lnode[i] + 1] - G.n2e_ptr[lnode[i]];
  for (int i = 0; i < nl; i++) n2e_ptr[i + 1] += n2e_ptr[i];
  const long long n2e_total = n2e_ptr[nl];
  n2e_koff.resize(n2e_total); n2e_roff.resize(n2e_total); n2e_nd.resize(n2e_total);
  n2e_fe.resize(n2e_total); n2e_loc.resize(n2e_total);
  // node-major storage of the element-tangent rows: slot u lives at KeN[u*chunk] (record models: no KeN, n2e_ksrc)
  chunk = ndf * cp_stride;
  kn_total = rec_mode ? 0 : n2e_total * chunk;
  n2e_ksrc.assign(rec_mode ? n2e_total : 0, 0);
  std::vector<Group>& WG = nparts == 1 ? groups : lgroups;   // the local groups being built
  for (auto& g : WG) g.kdst.assign(rec_mode ? 0 : (size_t)g.n() * ele_kind(g.kind).nen, 0);
  // receive-buffer layout: per source rank, chunks in (owned node ascending, FE order)
  std::vector<long long> rk(nparts, 0), rr(nparts, 0), rc(nparts, 0);
  std::vector<std::vector<long long>> in_kdst(nparts);
  for (int i = 0; i < nl; i++) if (owned[i]) {
    const int n = lnode[i];
    for (long long t = G.n2e_ptr[n], u = n2e_ptr[i]; t < G.n2e_ptr[n + 1]; t++, u++) {
      const long long e = G.n2e_fe[t];
      const EleKind* k; G.conn_of(e, &k);
      const int a = G.n2e_loc[t];
      const long long nd = k->nen * k->ndf;
      n2e_fe[u] = e; n2e_loc[u] = (uint8_t)a; n2e_nd[u] = (uint8_t)nd;
      n2e_koff[u] = u * chunk;
      const int src = part_fe[e];
      if (src == rank) {
        const int gi = G.fe_group[e];
        const long long l = nparts == 1 ? G.fe_local[e] : fe_local[g_fe_to_local[e]];
        if (rec_mode) n2e_ksrc[u] = brick_rec_desc(rec_off[gi] + l * kBrickRec, a);
        else WG[gi].kdst[(size_t)l * k->nen + a] = u * chunk;
        n2e_roff[u] = re_off[gi] + l * nd + (long long)a * k->ndf;
      } else {
        if (rec_mode) n2e_ksrc[u] = rk[src];   // relative to that peer's block of the receive buffer, fixed below
        else in_kdst[src].push_back(u * chunk);
        n2e_roff[u] = rr[src];      // relative to that peer's block, fixed below
        rk[src] += chunk; rr[src] += k->ndf; rc[src]++;
      }
    }
  }
  // send side: for every other rank s, the chunks of MY elements at nodes s owns, enumerated the
  // way s enumerates them (its owned nodes ascending, global FE order)
  std::vector<long long> sk(nparts, 0), sr(nparts, 0), sc(nparts, 0);
  struct Out { int gi; long long l; int a, nen; long long rsrc; };
  std::vector<std::vector<Out>> outs(nparts);
  if (nparts > 1) {
    for (int i = 0; i < nl; i++) {
      const int n = lnode[i];
      const int s = owner[n];
      if (s == rank) continue;
      for (long long t = G.n2e_ptr[n]; t < G.n2e_ptr[n + 1]; t++) {
        const long long e = G.n2e_fe[t];
        if (part_fe[e] != rank) continue;
        const EleKind* k; G.conn_of(e, &k);
        const int a = G.n2e_loc[t], gi = G.fe_group[e];
        const long long nd = k->nen * k->ndf, l = fe_local[g_fe_to_local[e]];
        outs[s].push_back({gi, l, a, k->nen, re_off[gi] + l * nd + (long long)a * k->ndf});
        sk[s] += chunk; sr[s] += k->ndf; sc[s]++;
      }
    }
  }
  peers.clear(); pr_src.clear(); pr_dst.clear(); uk_src.clear(); uk_dst.clear(); pk_src.clear();
  send_k_total = recv_k_total = send_r_total = recv_r_total = 0;
  std::vector<long long> recv_r_base(nparts, 0), recv_k_base(nparts, 0);
  for (int s = 0; s < nparts; s++) {
    if (s == rank || (sc[s] == 0 && rc[s] == 0)) continue;
    Peer p; p.rank = s;
    p.send_k = sk[s]; p.send_r = sr[s]; p.recv_k = rk[s]; p.recv_r = rr[s];
    p.chunks_out = sc[s]; p.chunks_in = rc[s];
    p.send_k_base = send_k_total; p.send_r_base = send_r_total; p.recv_k_base = recv_k_total; p.recv_r_base = recv_r_total;
    recv_r_base[s] = recv_r_total; recv_k_base[s] = recv_k_total;
    long long dk = send_k_total, dr = send_r_total;
    for (const Out& o : outs[s]) {
      // the element kernel writes straight into the send buffer; record models gather the rows out of the record
      // (pack_rows_rec_kernel: chunk c of the send buffer <- pk_src[c])
      if (rec_mode) pk_src.push_back(brick_rec_desc(rec_off[o.gi] + o.l * kBrickRec, o.a));
      else WG[o.gi].kdst[(size_t)o.l * o.nen + o.a] = -(dk + 1);
      pr_src.push_back(o.rsrc); pr_dst.push_back(dr);
      dk += chunk; dr += ndf;
    }
    long long rkpos = recv_k_total;
    for (long long d : in_kdst[s]) { uk_src.push_back(rkpos); uk_dst.push_back(d); rkpos += chunk; }
    send_k_total += sk[s]; send_r_total += sr[s]; recv_k_total += rk[s]; recv_r_total += rr[s];
    peers.push_back(p);
  }
  // remote residual slots: absolute offset into the receive buffer, encoded as -(x+1)
  if (nparts > 1)
    for (long long u = 0; u < n2e_total; u++) {
      const int src = part_fe[n2e_fe[u]];
      if (src != rank) {
        n2e_roff[u] = -(recv_r_base[src] + n2e_roff[u] + 1);
        if (rec_mode) n2e_ksrc[u] = dense_rows_desc(recv_k_base[src] + n2e_ksrc[u]);
      }
    }

  // ---- DOF graph -> sparse pattern of the owned rows.  Every free dof of node n is coupled
  // with every free dof of every node sharing an element with n (FE_Element::getID spans all
  // dofs of its nodes), so the rows/columns of one node share a single sorted list.  setSize()
  // then insertion-sorts diag + adjacency, i.e. the list including the dof itself. ----
  ncol_ptr.assign((size_t)nl + 1, 0);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
    for (int i = 0; i < nl; i++) {
      if (!owned[i]) continue;
      G.nbrs(lnode[i], tmp);
      long long c = 0;
      if (!have_mp) { for (int w : tmp) for (int j = 0; j < ndf; j++) if (gid[(size_t)w * ndf + j] >= 0) c++; }
      else {   // tied dofs put one equation under several neighbours: count distinct equations
        std::vector<int> q;
        for (int w : tmp) for (int j = 0; j < ndf; j++) if (gid[(size_t)w * ndf + j] >= 0) q.push_back(gid[(size_t)w * ndf + j]);
        std::sort(q.begin(), q.end());
        c = (long long)(std::unique(q.begin(), q.end()) - q.begin());
      }
      ncol_ptr[i + 1] = c;
    }
  }
  for (int i = 0; i < nl; i++) ncol_ptr[i + 1] += ncol_ptr[i];
  ncol.resize(ncol_ptr[nl]);
  long long too_long = 0;
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096) reduction(+ : too_long)
    for (int i = 0; i < nl; i++) {
      if (!owned[i]) continue;
      G.nbrs(lnode[i], tmp);
      int* out = &ncol[ncol_ptr[i]];
      long long c = 0;
      if (!have_mp) {
        for (int w : tmp) for (int j = 0; j < ndf; j++) { int q = gid[(size_t)w * ndf + j]; if (q >= 0) out[c++] = q; }
        std::sort(out, out + c);
      } else {
        std::vector<int> q;
        for (int w : tmp) for (int j = 0; j < ndf; j++) if (gid[(size_t)w * ndf + j] >= 0) q.push_back(gid[(size_t)w * ndf + j]);
        std::sort(q.begin(), q.end());
        q.erase(std::unique(q.begin(), q.end()), q.end());
        c = (long long)q.size();
        std::copy(q.begin(), q.end(), out);
      }
      if (c >= (have_mp ? 0x1FFF : 0xFFFF)) too_long++;   // with MP constraints 3 bits of a position carry the duplicate rank
    }
  }
  if (too_long) { err = "a node couples with more than 65534 equations (8190 with MP constraints)"; return XB_ERR_UNSUPPORTED; }

  // shared rows: the union of the column lists of every node with a dof on the equation
  row_of_dev = row_of;
  std::vector<std::vector<int>> irr_cols;
  irr_row.clear(); irr_own_ptr.assign(1, 0); irr_own.clear();
  if (have_mp) {
    std::vector<std::pair<int, int>> own;     // (row, node*ndf+dof), DOF_Group (node) order
    for (int i = 0; i < nl; i++) for (int j = 0; j < ndf; j++)
      if (owned[i] && gshared[(size_t)lnode[i] * ndf + j]) { own.push_back({row_of[(size_t)i * ndf + j], i * ndf + j}); row_of_dev[(size_t)i * ndf + j] = -1; }
    std::stable_sort(own.begin(), own.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    for (size_t u = 0; u < own.size(); u++) {
      if (u == 0 || own[u].first != own[u - 1].first) {
        if (u) irr_own_ptr.push_back((long long)irr_own.size());
        irr_row.push_back(own[u].first); irr_cols.emplace_back();
      }
      irr_own.push_back(own[u].second);
      const int i = own[u].second / ndf;
      std::vector<int>& cl = irr_cols.back();
      if (n2e_ptr[i + 1] == n2e_ptr[i]) cl.push_back(id[own[u].second]);
      else cl.insert(cl.end(), &ncol[ncol_ptr[i]], &ncol[ncol_ptr[i + 1]]);
    }
    if (!own.empty()) irr_own_ptr.push_back((long long)irr_own.size());
    for (auto& cl : irr_cols) { std::sort(cl.begin(), cl.end()); cl.erase(std::unique(cl.begin(), cl.end()), cl.end()); }
  }

  ptr.assign((size_t)nrows + 1, 0);
  max_row = 0;
  for (int i = 0; i < nl; i++) {
    if (!owned[i]) continue;
    long long L = ncol_ptr[i + 1] - ncol_ptr[i];
    bool isolated = n2e_ptr[i + 1] == n2e_ptr[i];
    // the assembly task of a node accumulates over the node's whole column list even when every one of its
    // rows is a shared (equalDOF) row: size the accumulator for the list, not for the rows it writes
    max_row = std::max<long long>(max_row, isolated ? 1 : L);
    for (int j = 0; j < ndf; j++) {
      int r = row_of_dev[(size_t)i * ndf + j];
      if (r >= 0) ptr[r + 1] = isolated ? 1 : L;
    }
  }
  irr_max_row = 0;
  for (size_t w = 0; w < irr_row.size(); w++) {
    ptr[irr_row[w] + 1] = (long long)irr_cols[w].size();
    irr_max_row = std::max<int>(irr_max_row, (int)irr_cols[w].size());
    if (irr_cols[w].size() >= 0x1FFF) { err = "a shared equation couples with more than 8190 equations"; return XB_ERR_UNSUPPORTED; }
  }
  for (int r = 0; r < nrows; r++) ptr[r + 1] += ptr[r];
  const long long nz = ptr[nrows];
  idx.resize(nz);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int i = 0; i < nl; i++) {
    if (!owned[i]) continue;
    long long L = ncol_ptr[i + 1] - ncol_ptr[i];
    bool isolated = n2e_ptr[i + 1] == n2e_ptr[i];
    for (int j = 0; j < ndf; j++) {
      int r = row_of_dev[(size_t)i * ndf + j];
      if (r < 0) continue;
      if (isolated) idx[ptr[r]] = id[(size_t)i * ndf + j];
      else std::copy(&ncol[ncol_ptr[i]], &ncol[ncol_ptr[i]] + L, &idx[ptr[r]]);
    }
  }
  for (size_t w = 0; w < irr_row.size(); w++) std::copy(irr_cols[w].begin(), irr_cols[w].end(), &idx[ptr[irr_row[w]]]);

  // ---- BandGeneral / ProfileSPD: where every pattern entry (column r, row idx[k]) lives in the SOE's own array ----
  a_loc.clear(); profile_diag.clear(); a_total = nz; band_sub = band_super = 0;
  if (soe_store == XB_SOE_BAND_GEN) {
    // BandGenLinSOE::setSize (BandGenLinSOE.cpp:116-147): the largest vertex - other and other - vertex over the DOF graph
    for (int r = 0; r < nrows; r++)
      for (long long k = ptr[r]; k < ptr[r + 1]; k++) {
        const int diff = r - idx[k];
        if (diff > band_super) band_super = diff;
        if (-diff > band_sub) band_sub = -diff;
      }
    const long long ldA = 2LL * band_sub + band_super + 1;
    a_total = ldA * nrows;
    a_loc.resize(nz);
    // addA (BandGenLinSOE.cpp:208-249): column col, row row -> A[col ldA + numSubD + numSuperD - (col - row)]
    for (int col = 0; col < nrows; col++)
      for (long long k = ptr[col]; k < ptr[col + 1]; k++) a_loc[k] = col * ldA + band_sub + band_super - (col - idx[k]);
  } else if (soe_store == XB_SOE_PROFILE_SPD) {
    // ProfileSPDLinSOE::setSize (ProfileSPDLinSOE.cpp:115-168): column heights, then running sums (1-based)
    profile_diag.assign(nrows, 0);
    for (int r = 0; r < nrows; r++)
      for (long long k = ptr[r]; k < ptr[r + 1]; k++) profile_diag[r] = std::max(profile_diag[r], r - idx[k]);
    if (nrows > 0) profile_diag[0] = 1;
    for (int j = 1; j < nrows; j++) profile_diag[j] = profile_diag[j] + 1 + profile_diag[j - 1];
    a_total = nrows > 0 ? profile_diag[nrows - 1] : 0;
    a_loc.assign(nz, -1);
    // addA (ProfileSPDLinSOE.cpp:214-243): row <= col inside the profile -> A[iDiagLoc[col] - 1 + row - col]
    for (int col = 0; col < nrows; col++)
      for (long long k = ptr[col]; k < ptr[col + 1]; k++)
        if (idx[k] <= col) a_loc[k] = (long long)profile_diag[col] - 1 + (idx[k] - col);
  }

  // position of every owned dof's own equation in its node's list (the diagonal entry of A)
  diagpos.assign((size_t)nl * ndf, 0xFFFF);
  for (int i = 0; i < nl; i++) {
    if (!owned[i]) continue;
    const int* cols = &ncol[ncol_ptr[i]];
    const long long L = ncol_ptr[i + 1] - ncol_ptr[i];
    const bool isolated = n2e_ptr[i + 1] == n2e_ptr[i];
    for (int j = 0; j < ndf; j++) {
      const int q = id[(size_t)i * ndf + j];
      if (q < 0 || row_of_dev[(size_t)i * ndf + j] < 0) continue;    // shared rows: irr_diag
      diagpos[(size_t)i * ndf + j] = isolated ? 0 : (uint16_t)(std::lower_bound(cols, cols + L, q) - cols);
    }
  }

  // ---- per (owned node, adjacent element) positions of the element's dofs in the node's list ----
  colpos.assign((size_t)n2e_total * cp_stride, 0xFFFF);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int i = 0; i < nl; i++) {
    if (!owned[i]) continue;
    const int* cols = &ncol[ncol_ptr[i]];
    const long long L = ncol_ptr[i + 1] - ncol_ptr[i];
    for (long long t = n2e_ptr[i]; t < n2e_ptr[i + 1]; t++) {
      const EleKind* k; const int* c = G.conn_of(n2e_fe[t], &k);
      uint16_t* cp = &colpos[(size_t)t * cp_stride];
      for (int a = 0; a < k->nen; a++)
        for (int j = 0; j < k->ndf; j++) {
          int q = gid[(size_t)c[a] * ndf + j];
          if (q < 0) continue;
          const int* it = std::lower_bound(cols, cols + L, q);
          cp[a * k->ndf + j] = (uint16_t)(it - cols);
        }
      if (have_mp) dup_ranks(cp, k->nen * k->ndf);
    }
  }
  // shared rows: their contributions in (FE_Element, element dof) order -- the order addA / addB see them
  irr_ptr.assign(1, 0); irr_src.clear(); irr_roff.clear(); irr_cp.clear(); irr_diag.clear();
  for (size_t w = 0; w < irr_row.size(); w++) {
    const std::vector<int>& cl = irr_cols[w];
    struct Ent { long long fe; int dof; long long t; int j; };
    std::vector<Ent> ents;
    for (long long o = irr_own_ptr[w]; o < irr_own_ptr[w + 1]; o++) {
      const int i = irr_own[o] / ndf, j = irr_own[o] % ndf;
      for (long long t = n2e_ptr[i]; t < n2e_ptr[i + 1]; t++) ents.push_back({n2e_fe[t], n2e_loc[t] * ndf + j, t, j});
    }
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.fe != b.fe ? a.fe < b.fe : a.dof < b.dof; });
    for (const Ent& en : ents) {
      const EleKind* k; const int* c = G.conn_of(en.fe, &k);
      // the element-matrix row: dense in KeN, or -- record models -- the slot's descriptor and the dof (descriptor * 4 + dof)
      irr_src.push_back(rec_mode ? n2e_ksrc[en.t] * 4 + en.j : en.t * chunk + (long long)en.j * cp_stride);
      irr_roff.push_back(n2e_roff[en.t] >= 0 ? n2e_roff[en.t] + en.j : n2e_roff[en.t] - en.j);   // < 0: -(offset in recvR + 1)
      const size_t base = irr_cp.size();
      irr_cp.resize(base + cp_stride, 0xFFFF);
      for (int a = 0; a < k->nen; a++)
        for (int j = 0; j < k->ndf; j++) {
          const int q = gid[(size_t)c[a] * ndf + j];
          if (q < 0) continue;
          irr_cp[base + a * k->ndf + j] = (uint16_t)(std::lower_bound(cl.begin(), cl.end(), q) - cl.begin());
        }
      dup_ranks(&irr_cp[base], k->nen * k->ndf);
    }
    irr_ptr.push_back((long long)irr_src.size());
    irr_diag.push_back((uint16_t)(std::lower_bound(cl.begin(), cl.end(), row_geq[irr_row[w]]) - cl.begin()));
  }

  // 32-bit descriptors for the hand-tuned record assembly (plain brick models): offset in units of 36 doubles << 4 |
  // local node, or | 8 for dense rows in the receive buffer, which lies right behind the records
  n2e_ksrc32.clear();
  {
    long long most = 0;
    for (int i = 0; i < nl; i++) most = std::max(most, n2e_ptr[i + 1] - n2e_ptr[i]);
    fast_asm_ok = rec_mode && !have_mp && max_row <= 96 && most <= 32 && (rec_total + recv_k_total) / 36 < (1ll << 28);
  }
  if (fast_asm_ok) {
    n2e_ksrc32.resize(n2e_total);
    for (long long u = 0; u < n2e_total; u++) {
      const long long d = n2e_ksrc[u];
      n2e_ksrc32[u] = (d & 1) ? (uint32_t)((((d >> 4) / 36) << 4) | ((d >> 1) & 7)) : (uint32_t)((((rec_total + (d >> 4)) / 36) << 4) | 8);
    }
  }

  max_dup = 0;
  if (have_mp) {
    for (uint16_t v : colpos) if (v != 0xFFFF) max_dup = std::max(max_dup, (int)(v >> 13));
    for (uint16_t v : irr_cp) if (v != 0xFFFF) max_dup = std::max(max_dup, (int)(v >> 13));
    // 0xFFFF is "no column"; a real position never reaches 0x1FFF here, so rank 7 + position 0x1FFF cannot occur
  }

  // ---- node order for the ranged formTangent (single-batch models; else one range) ----
  {
    const std::vector<Group>& FG = nparts == 1 ? groups : lgroups;
    // With a host destination, xb_form_tangent sends the rows of range c to the host while range c+1 is still
    // being formed (chunk_a_ptr below); the element kernel and the assembly of the previous range also share the SMs.
    // one range per ~0.5 M elements, at most `want_ranges` (8): 8 at 4 M elements on one GPU, 1 on each of 8 GPUs --
    // ranges of less than that are launch-bound (N = 8 measured 2.16 ms per step with 8 ranges of 64 k elements)
    const int want = (int)std::max<long long>(1, std::min<long long>(want_ranges, ne / 500000));
    nchunk = (FG.size() == 1 && ne >= 65536 && !have_mp) ? want : 1;
    const long long per = nchunk > 1 ? (ne + nchunk - 1) / nchunk : ne;
    std::vector<int> ready(nl, -1);
    for (int i = 0; i < nl; i++) {
      if (!owned[i]) continue;
      int rdy = 0;
      for (long long t = n2e_ptr[i]; t < n2e_ptr[i + 1]; t++) {
        const long long ge = n2e_fe[t];
        if (part_fe[ge] != rank) { rdy = nchunk; break; }          // needs the interface exchange
        const long long le = nparts == 1 ? ge : g_fe_to_local[ge];  // local FE index == index in the batch
        const long long l = fe_local[le];
        rdy = std::max(rdy, nchunk > 1 ? (int)(l / per) : 0);
      }
      ready[i] = rdy;
    }
    chunk_node_ptr.assign((size_t)nchunk + 2, 0);
    for (int i = 0; i < nl; i++) if (ready[i] >= 0) chunk_node_ptr[ready[i] + 1]++;
    for (int c = 0; c <= nchunk; c++) chunk_node_ptr[c + 1] += chunk_node_ptr[c];
    node_perm.resize(chunk_node_ptr[nchunk + 1]);
    std::vector<long long> fill(chunk_node_ptr.begin(), chunk_node_ptr.end() - 1);
    for (int i = 0; i < nl; i++) if (ready[i] >= 0) node_perm[fill[ready[i]]++] = i;
    if (rec_mode) {
      // Morton order of the node coordinates inside every range (one cell size for all axes, 16 bits each): the
      // eight nodes of an element are assembled close together in time, so its record is read from HBM once
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int i = 0; i < nl; i++)
        for (int d = 0; d < ndm; d++) { lo[d] = std::min(lo[d], crd[(size_t)i * ndm + d]); hi[d] = std::max(hi[d], crd[(size_t)i * ndm + d]); }
      double ext = 0.0;
      for (int d = 0; d < ndm; d++) ext = std::max(ext, hi[d] - lo[d]);
      const double inv = ext > 0.0 ? 65535.0 / ext : 0.0;
      auto spread = [](uint64_t v) {   // 16 bits -> every third bit
        v &= 0xFFFF;
        v = (v | (v << 32)) & 0x1F00000000FFFFull;
        v = (v | (v << 16)) & 0x1F0000FF0000FFull;
        v = (v | (v << 8)) & 0x100F00F00F00F00Full;
        v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
        v = (v | (v << 2)) & 0x1249249249249249ull;
        return v;
      };
      std::vector<std::pair<uint64_t, int>> keyed(node_perm.size());
#pragma omp parallel for schedule(static)
      for (long long u = 0; u < (long long)node_perm.size(); u++) {
        const int i = node_perm[u];
        uint64_t key = 0;
        for (int d = 0; d < ndm; d++) key |= spread((uint64_t)((crd[(size_t)i * ndm + d] - lo[d]) * inv)) << d;
        keyed[u] = {key, i};
      }
      for (int c = 0; c <= nchunk; c++) std::sort(keyed.begin() + chunk_node_ptr[c], keyed.begin() + chunk_node_ptr[c + 1]);
      for (size_t u = 0; u < node_perm.size(); u++) node_perm[u] = keyed[u].second;
    }
    // what the assembly warp of node_perm[u] needs, in one record (assemble_A_kernel)
    const int TW = 3 + ndf;
    asm_task.assign((size_t)node_perm.size() * TW, 0);
    for (size_t u = 0; u < node_perm.size(); u++) {
      const int i = node_perm[u];
      long long* tk = &asm_task[u * TW];
      const long long ns = n2e_ptr[i + 1] - n2e_ptr[i];
      long long L = ncol_ptr[i + 1] - ncol_ptr[i];
      if (ns == 0) L = 1;   // a node with no element: its rows hold the (zero) diagonal only
      tk[0] = n2e_ptr[i]; tk[1] = ns | (L << 32); tk[2] = i;
      for (int j = 0; j < ndf; j++) {
        const int r = row_of_dev[(size_t)i * ndf + j];
        tk[3 + j] = r >= 0 ? ptr[r] : -1;
      }
    }
    // rows each range completes: streamable when range c owns exactly the rows [r_c, r_{c+1}) of A
    const int ng = nchunk;
    auto group_of = [&](int c) { return c; };
    chunk_a_ptr.assign((size_t)ng + 2, 0);
    rows_streamable = nchunk > 1 && a_loc.empty();
    int next_row = 0;
    for (int gI = 0, c = 0; gI <= ng && rows_streamable; gI++) {
      long long cnt = 0; int lo = nrows, hi = -1;
      for (; c <= nchunk && group_of(c) == gI; c++)
        for (long long u = chunk_node_ptr[c]; u < chunk_node_ptr[c + 1]; u++)
          for (int j = 0; j < ndf; j++) {
            const int r = row_of_dev[(size_t)node_perm[u] * ndf + j];
            if (r < 0) continue;
            cnt++; lo = std::min(lo, r); hi = std::max(hi, r);
          }
      if (cnt && (lo != next_row || hi - lo + 1 != cnt)) rows_streamable = false;
      if (cnt) next_row = hi + 1;
      chunk_a_ptr[gI + 1] = ptr[next_row];
    }
    if (next_row != nrows) rows_streamable = false;
  }

  // ---- commit the local element groups ----
  if (nparts > 1) groups.swap(lgroups);
  for (size_t gi = 0; gi < groups.size(); gi++) { groups[gi].ke_off = ke_off[gi]; groups[gi].rec_off = rec_off[gi]; groups[gi].re_off = re_off[gi]; groups[gi].gp_off = gp_off[gi]; }
  is_setup = true;
  return neq;
}

int HostModel::scatter_map(long long e0, long long e1, long long* map) const {
  if (!is_setup || e0 < 0 || e1 > ne || e0 > e1) return XB_ERR_ARG;
  for (long long e = e0; e < e1; e++) {
    const Group& g = groups[fe_group[e]];
    const EleKind& k = ele_kind(g.kind);
    const int nd = k.nen * k.ndf;
    const int* c = &g.conn[(size_t)fe_local[e] * k.nen];   // local node indices
    long long* out = map + (e - e0) * nd * nd;
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++) {
        // CSR: entry (i,j) -> row id(i), column id(j).  CSC: entry (i,j) -> column id(j), row id(i).
        // The location is where the assembly kernels put it: the owned row of the major equation, at the
        // position of the minor equation in that row's (sorted) column list.
        const int majordof = soe_kind == XB_SOE_SPARSE_GEN_ROW ? i : j;
        const int minordof = soe_kind == XB_SOE_SPARSE_GEN_ROW ? j : i;
        const int qmaj = id[(size_t)c[majordof / k.ndf] * ndf + majordof % k.ndf];
        const int qmin = id[(size_t)c[minordof / k.ndf] * ndf + minordof % k.ndf];
        out[i * nd + j] = -1;
        if (qmaj < 0 || qmin < 0) continue;
        const auto rit = std::lower_bound(row_geq.begin(), row_geq.end(), qmaj);
        if (rit == row_geq.end() || *rit != qmaj) continue;                     // a row another rank owns
        const long long r = rit - row_geq.begin();
        const int* b = &idx[ptr[r]]; const int* en = &idx[ptr[r + 1]];
        const int* it = std::lower_bound(b, en, qmin);
        if (it != en && *it == qmin) {
          const long long k = ptr[r] + (it - b);
          out[i * nd + j] = a_loc.empty() ? k : a_loc[k];
        }
      }
  }
  return XB_OK;
}

}  // namespace xb
