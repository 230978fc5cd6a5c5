// Host-side analysis set-up for the device path (see host_model.hpp for the
// reference file:line each step mirrors).
#include "host_model.hpp"

#include <algorithm>
#include <cstring>
#include <numeric>

#include "../../include/xara_b200.h"

namespace xb {

static const EleKind kBrick{8, 3, 8, 6, 3};
static const EleKind kQuad{4, 2, 4, 3, 3};  // par kept: thickness, b1, b2

const EleKind& ele_kind(int kind) { return kind == XB_ELE_STDBRICK ? kBrick : kQuad; }

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

int HostModel::add_elements(int kind, int n, const int* tags, const int* conn, const int* mat_tags,
                            const double* par, int par_stride) {
  if (is_setup) { err = "xb_add_elements after xb_setup"; return XB_ERR_STATE; }
  if (kind != XB_ELE_STDBRICK && kind != XB_ELE_FOURNODEQUAD) { err = "xb_add_elements: unknown kind"; return XB_ERR_ARG; }
  const EleKind& k = ele_kind(kind);
  if (k.ndf != ndf) { err = "xb_add_elements: element dofs per node differ from the model's ndf"; return XB_ERR_UNSUPPORTED; }
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
      if ((int)p[1] != 0) { err = "FourNodeQuad: only PlaneStrain on the device path"; return XB_ERR_UNSUPPORTED; }
      if (p[2] != 0.0) { err = "FourNodeQuad: surface pressure not on the device path"; return XB_ERR_UNSUPPORTED; }
      q[0] = p[0]; q[1] = p[4]; q[2] = p[5];
    }
  }
  g.mat_kind = mk < 0 ? 0 : mk;
  if (n > 0) groups.push_back(std::move(g));
  return XB_OK;
}

int HostModel::add_loads(int n, const int* tags, const double* vals) {
  if (is_setup) { err = "xb_add_nodal_loads after xb_setup"; return XB_ERR_STATE; }
  load_node.insert(load_node.end(), tags, tags + n);
  load_val.insert(load_val.end(), vals, vals + (size_t)n * ndf);
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

int HostModel::setup(int numberer_, int soe_kind_) {
  if (is_setup) { err = "xb_setup called twice"; return XB_ERR_STATE; }
  if (numberer_ != XB_NUMBERER_PLAIN && numberer_ != XB_NUMBERER_RCM) { err = "unknown numberer"; return XB_ERR_ARG; }
  if (soe_kind_ != XB_SOE_SPARSE_GEN_COL && soe_kind_ != XB_SOE_SPARSE_GEN_ROW) { err = "unknown SOE kind"; return XB_ERR_ARG; }
  numberer = numberer_; soe_kind = soe_kind_;
  const int n_nodes = (int)node_tag.size();

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

  // ---- element connectivity: tags -> node indices ----
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
  id.assign((size_t)n_nodes * ndf, -2);
  for (size_t i = 0; i < sp_node.size(); i++) {
    int ix = nidx(sp_node[i]);
    if (ix < 0) { err = "fix references an unknown node tag"; return XB_ERR_ARG; }
    id[(size_t)ix * ndf + sp_dof[i]] = -1;
  }
  load.assign((size_t)n_nodes * ndf, 0.0);
  for (size_t i = 0; i < load_node.size(); i++) {
    int ix = nidx(load_node[i]);
    if (ix < 0) { err = "load references an unknown node tag"; return XB_ERR_ARG; }
    for (int j = 0; j < ndf; j++) load[(size_t)ix * ndf + j] += load_val[i * ndf + j];
  }

  // ---- FE_Element order: Domain element map by ascending tag (PlainHandler.cpp:228-250) ----
  ne = 0;
  for (auto& g : groups) ne += g.n();
  fe_group.resize(ne); fe_local.resize(ne);
  {
    bool simple = groups.size() == 1 && std::is_sorted(groups[0].tag.begin(), groups[0].tag.end());
    if (simple) {
      for (long long e = 0; e < ne; e++) { fe_group[e] = 0; fe_local[e] = (int)e; }
    } else {
      std::vector<long long> order(ne);
      std::vector<int> tg(ne), gg(ne), ll(ne);
      long long c = 0;
      for (size_t gi = 0; gi < groups.size(); gi++)
        for (long long l = 0; l < groups[gi].n(); l++) { tg[c] = groups[gi].tag[l]; gg[c] = (int)gi; ll[c] = (int)l; c++; }
      std::iota(order.begin(), order.end(), 0LL);
      std::sort(order.begin(), order.end(), [&](long long a, long long b) { return tg[a] < tg[b]; });
      for (long long e = 0; e < ne; e++) { fe_group[e] = gg[order[e]]; fe_local[e] = ll[order[e]]; }
      for (long long e = 1; e < ne; e++) if (tg[order[e]] == tg[order[e - 1]]) { err = "duplicate element tag"; return XB_ERR_ARG; }
    }
    for (auto& g : groups) for (size_t i = 1; i < g.tag.size() && groups.size() == 1; i++)
      if (g.tag[i] == g.tag[i - 1]) { err = "duplicate element tag"; return XB_ERR_ARG; }
  }
  ke_total = re_total = ngp = 0;
  for (auto& g : groups) {
    const EleKind& k = ele_kind(g.kind);
    const long long nd = k.nen * k.ndf;
    g.ke_off = ke_total; g.re_off = re_total; g.gp_off = ngp;
    ke_total += g.n() * nd * nd; re_total += g.n() * nd; ngp += g.n() * k.nip;
  }

  // ---- node -> FE elements (FE order) ----
  n2e_ptr.assign((size_t)n_nodes + 1, 0);
  for (long long e = 0; e < ne; e++) {
    const Group& g = groups[fe_group[e]];
    const EleKind& k = ele_kind(g.kind);
    const int* c = &g.conn[(size_t)fe_local[e] * k.nen];
    for (int a = 0; a < k.nen; a++) n2e_ptr[c[a] + 1]++;
  }
  for (int n = 0; n < n_nodes; n++) n2e_ptr[n + 1] += n2e_ptr[n];
  const long long n2e_total = n2e_ptr[n_nodes];
  n2e_koff.resize(n2e_total); n2e_roff.resize(n2e_total); n2e_nd.resize(n2e_total);
  n2e_fe.resize(n2e_total); n2e_loc.resize(n2e_total);
  {
    std::vector<long long> fill(n2e_ptr.begin(), n2e_ptr.end() - 1);
    for (long long e = 0; e < ne; e++) {
      const Group& g = groups[fe_group[e]];
      const EleKind& k = ele_kind(g.kind);
      const long long nd = k.nen * k.ndf, l = fe_local[e];
      const int* c = &g.conn[(size_t)l * k.nen];
      for (int a = 0; a < k.nen; a++) {
        long long t = fill[c[a]]++;
        n2e_fe[t] = (int)e; n2e_loc[t] = (uint8_t)a; n2e_nd[t] = (uint8_t)nd;
        n2e_koff[t] = g.ke_off + l * nd * nd + (long long)a * k.ndf * nd;
        n2e_roff[t] = g.re_off + l * nd + (long long)a * k.ndf;
      }
    }
  }

  // neighbours of node n (sorted, unique, including n when it has an element)
  auto collect_nbrs = [&](int n, std::vector<int>& out) {
    out.clear();
    for (long long t = n2e_ptr[n]; t < n2e_ptr[n + 1]; t++) {
      const Group& g = groups[fe_group[n2e_fe[t]]];
      const EleKind& k = ele_kind(g.kind);
      const int* c = &g.conn[(size_t)fe_local[n2e_fe[t]] * k.nen];
      out.insert(out.end(), c, c + k.nen);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
  };

  // ---- numbering ----
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
      for (int n = 0; n < n_nodes; n++) { collect_nbrs(n, tmp); nb_ptr[n + 1] = (long long)tmp.size(); }
    }
    for (int n = 0; n < n_nodes; n++) nb_ptr[n + 1] += nb_ptr[n];
    std::vector<int> nb(nb_ptr[n_nodes]);
#pragma omp parallel
    {
      std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
      for (int n = 0; n < n_nodes; n++) { collect_nbrs(n, tmp); std::copy(tmp.begin(), tmp.end(), nb.begin() + nb_ptr[n]); }
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
    for (int j = 0; j < ndf; j++) if (id[(size_t)n * ndf + j] == -2) id[(size_t)n * ndf + j] = eqn++;
  }
  neq = eqn;

  // ---- DOF graph -> sparse pattern.  Every free dof of node n is coupled with every free
  // dof of every node sharing an element with n (FE_Element::getID spans all dofs of its
  // nodes), so the rows/columns of one node share a single sorted list.  setSize() then
  // insertion-sorts diag + adjacency, i.e. the list including the dof itself. ----
  ncol_ptr.assign((size_t)n_nodes + 1, 0);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
    for (int n = 0; n < n_nodes; n++) {
      collect_nbrs(n, tmp);
      long long c = 0;
      for (int w : tmp) for (int j = 0; j < ndf; j++) if (id[(size_t)w * ndf + j] >= 0) c++;
      ncol_ptr[n + 1] = c;
    }
  }
  for (int n = 0; n < n_nodes; n++) ncol_ptr[n + 1] += ncol_ptr[n];
  ncol.resize(ncol_ptr[n_nodes]);
  long long too_long = 0;
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096) reduction(+ : too_long)
    for (int n = 0; n < n_nodes; n++) {
      collect_nbrs(n, tmp);
      int* out = &ncol[ncol_ptr[n]];
      long long c = 0;
      for (int w : tmp) for (int j = 0; j < ndf; j++) { int q = id[(size_t)w * ndf + j]; if (q >= 0) out[c++] = q; }
      std::sort(out, out + c);
      if (c >= 0xFFFF) too_long++;
    }
  }
  if (too_long) { err = "a node couples with more than 65534 equations"; return XB_ERR_UNSUPPORTED; }

  ptr.assign((size_t)neq + 1, 0);
  max_row = 0;
  for (int n = 0; n < n_nodes; n++) {
    long long L = ncol_ptr[n + 1] - ncol_ptr[n];
    bool isolated = n2e_ptr[n + 1] == n2e_ptr[n];
    for (int j = 0; j < ndf; j++) {
      int r = id[(size_t)n * ndf + j];
      if (r >= 0) { ptr[r + 1] = isolated ? 1 : L; max_row = std::max<long long>(max_row, ptr[r + 1]); }
    }
  }
  for (int r = 0; r < neq; r++) ptr[r + 1] += ptr[r];
  const long long nz = ptr[neq];
  if (nz > 0x7fffffffLL * 4) { err = "pattern too large"; return XB_ERR_UNSUPPORTED; }
  idx.resize(nz);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int n = 0; n < n_nodes; n++) {
    long long L = ncol_ptr[n + 1] - ncol_ptr[n];
    bool isolated = n2e_ptr[n + 1] == n2e_ptr[n];
    for (int j = 0; j < ndf; j++) {
      int r = id[(size_t)n * ndf + j];
      if (r < 0) continue;
      if (isolated) idx[ptr[r]] = r;
      else std::copy(&ncol[ncol_ptr[n]], &ncol[ncol_ptr[n]] + L, &idx[ptr[r]]);
    }
  }

  // ---- per (node, adjacent element) positions of the element's dofs in the node's list ----
  colpos.assign((size_t)n2e_total * cp_stride, 0xFFFF);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int n = 0; n < n_nodes; n++) {
    const int* cols = &ncol[ncol_ptr[n]];
    const long long L = ncol_ptr[n + 1] - ncol_ptr[n];
    for (long long t = n2e_ptr[n]; t < n2e_ptr[n + 1]; t++) {
      const Group& g = groups[fe_group[n2e_fe[t]]];
      const EleKind& k = ele_kind(g.kind);
      const int* c = &g.conn[(size_t)fe_local[n2e_fe[t]] * k.nen];
      uint16_t* cp = &colpos[(size_t)t * cp_stride];
      for (int a = 0; a < k.nen; a++)
        for (int j = 0; j < k.ndf; j++) {
          int q = id[(size_t)c[a] * ndf + j];
          if (q < 0) continue;
          const int* it = std::lower_bound(cols, cols + L, q);
          cp[a * k.ndf + j] = (uint16_t)(it - cols);
        }
    }
  }
  is_setup = true;
  return neq;
}

int HostModel::scatter_map(long long e0, long long e1, long long* map) const {
  if (!is_setup || e0 < 0 || e1 > ne || e0 > e1) return XB_ERR_ARG;
  for (long long e = e0; e < e1; e++) {
    const Group& g = groups[fe_group[e]];
    const EleKind& k = ele_kind(g.kind);
    const int nd = k.nen * k.ndf;
    const int* c = &g.conn[(size_t)fe_local[e] * k.nen];
    long long* out = map + (e - e0) * nd * nd;
    // the n2e slot of (node c[a], element e)
    std::vector<long long> slot(k.nen);
    for (int a = 0; a < k.nen; a++) {
      slot[a] = -1;
      for (long long t = n2e_ptr[c[a]]; t < n2e_ptr[c[a] + 1]; t++)
        if (n2e_fe[t] == e && n2e_loc[t] == a) { slot[a] = t; break; }
    }
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++) {
        // CSR: entry (i,j) -> row id(i), column id(j).  CSC: entry (i,j) -> column id(j), row id(i).
        int owner = soe_kind == XB_SOE_SPARSE_GEN_ROW ? i : j;
        int other = soe_kind == XB_SOE_SPARSE_GEN_ROW ? j : i;
        int ro = id[(size_t)c[owner / k.ndf] * ndf + owner % k.ndf];
        uint16_t cp = colpos[(size_t)slot[owner / k.ndf] * cp_stride + other];
        out[i * nd + j] = (ro < 0 || cp == 0xFFFF) ? -1 : ptr[ro] + cp;
      }
  }
  return XB_OK;
}

}  // namespace xb
