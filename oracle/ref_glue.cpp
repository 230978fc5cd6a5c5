// TEST INFRASTRUCTURE ONLY -- the reference-side binding of INTEGRATION.md, built for real.
//
// The reference's OWN analysis objects (AnalysisModel, PlainHandler, numberer, SparseGenCol/Row
// LinearSOE + solver, NewtonRaphson, CTestNormDispIncr, LoadControl::newStep) run a static analysis
// in which only the three loops of the hot path are replaced by calls through the C ABI of
// include/xara_b200.h:
//     IncrementalIntegrator::formTangent   -> xb_form_tangent   (A lands in the SOE's own array)
//     StaticIntegrator::formUnbalance      -> xb_apply_load + xb_form_unbalance
//     LoadControl::update -> updateDomain  -> xb_incr_trial_disp + xb_update
//     IncrementalIntegrator::commit        -> xb_commit (+ the reference's own commitDomain for the nodes)
// The device model is read OUT OF THE REFERENCE'S Domain (nodes, SP constraints, Brick / FourNodeQuad
// elements and their J2Plasticity / ElasticIsotropic materials, nodal loads): no script or model
// command changes.  tests/test_gpu_parity.py compares the iteration counts and norms of this run
// with those of the unmodified reference run -- the reference's own convergence test decides both.
//
// Built by oracle/ref_build.mk into oracle/_ref/libref_glue.so (links libxara_ref.a and
// xara_b200/libxara_b200.so).  It re-uses the harness (model building through the reference's
// classes) by including its translation unit; private members the glue has to read (element
// connectivity, material parameters -- a maintainer would add accessors) are opened for this
// translation unit only.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include <LoadControl.h>               // before the harness: it drops the defines half way through its own includes
#include <Newmark.h>
#include <DisplacementControl.h>
#include <Node.h>
#include <Element.h>
#include <ForceBeamColumn2d.h>
#include <ForceBeamColumn3d.h>
#include <FiberSection2d.h>
#include <FiberSection3d.h>
#include <Steel02.h>
#include <Steel01.h>
#include <Concrete01.h>
#include <ElasticPPMaterial.h>
#include <SectionAggregator.h>
#include <Concrete02.h>
#include <ElasticMaterial.h>
#include <LinearCrdTransf2d.h>
#include <LinearCrdTransf3d.h>
#include <PDeltaCrdTransf2d.h>
#include <CorotCrdTransf2d.h>
#include <TransformationDOF_Group.h>
#include <TransformationConstraintHandler.h>
#include <PDeltaCrdTransf3d.h>
#include <LobattoBeamIntegration.h>
#include <Brick.h>
#include <FourNodeQuad.h>
#include <J2Plasticity.h>
#include <ElasticIsotropicMaterial.h>
#include "ref_harness.cpp"
#undef private
#undef protected
#include <SP_ConstraintIter.h>
#include <MP_Constraint.h>
#include <MP_ConstraintIter.h>
#include <LoadPatternIter.h>
#include <NodalLoadIter.h>
#include <ElementalLoadIter.h>
#include <LinearSeries.h>

#include "../include/xara_b200.h"

namespace {

struct Glue {
  xb_model* x = nullptr;
  std::string err;
  int numberer = -1, soeKind = -1;
  std::map<int, int> pattern_state;    // load pattern tag -> 0 live at the last set-up, 1 frozen
};
std::map<void*, Glue> g_glue;

// A second `analysis` on a Domain that already lives on the device (configs[0]: gravity under LoadControl, `loadConst
// -time 0`, then the pushover under DisplacementControl): the element state stays where it is; the load patterns that
// were frozen since the last set-up become the device's constant loads, the new patterns its reference loads.
int reattach_xb(RefModel* m, int numberer, int soeKind, Glue& G) {
  if (numberer != G.numberer || soeKind != G.soeKind) { G.err = "glue: a later analysis must keep the numberer and the system of the first"; return -12; }
  Domain* dom = m->domain;
  bool froze = false;
  std::vector<int> nt; std::vector<double> nv;
  { LoadPatternIter& pi = dom->getLoadPatterns(); LoadPattern* lp;
    while ((lp = pi()) != nullptr) {
      const int tag = lp->getTag();
      const auto it = G.pattern_state.find(tag);
      if (it != G.pattern_state.end()) {
        if (it->second == 0 && lp->isConstant) { froze = true; it->second = 1; }
        else if (it->second == 0) { G.err = "glue: a live load pattern carried into a later analysis: freeze it with loadConst"; return -12; }
        continue;
      }
      LinearSeries* ls = dynamic_cast<LinearSeries*>(lp->theSeries);
      if (!ls || ls->cFactor != 1.0 || lp->isConstant) { G.err = "glue: new load pattern that is not a live Linear series: outside the device path"; return -12; }
      { ElementalLoadIter& eli = lp->getElementalLoads(); if (eli() != nullptr) { G.err = "glue: element loads in a pattern added after the set-up"; return -12; } }
      NodalLoadIter& li = lp->getNodalLoads(); NodalLoad* nl;
      while ((nl = li()) != nullptr) {
        int type; const Vector& v = nl->getData(type);
        nt.push_back(nl->getNodeTag());
        for (int d = 0; d < m->ndf; d++) nv.push_back(d < v.Size() ? v(d) : 0.0);
      }
      G.pattern_state[tag] = 0;
    } }
  if (froze && xb_load_const(G.x) < 0) { G.err = xb_last_error(); return -12; }
  if (xb_apply_load(G.x, dom->getCurrentTime()) < 0) { G.err = xb_last_error(); return -12; }       // loadConst -time t
  if (!nt.empty() && xb_set_nodal_loads(G.x, (int)nt.size(), nt.data(), nv.data()) < 0) { G.err = xb_last_error(); return -12; }
  return 0;
}

// INTEGRATION.md "B200Assembler::domainChanged": the Domain -> xb_model
int domain_to_xb(RefModel* m, int numberer, int soeKind, int device, Glue& G) {
  Domain* dom = m->domain;
  if (G.x) return reattach_xb(m, numberer, soeKind, G);
  G.numberer = numberer; G.soeKind = soeKind;
  xb_model* x = xb_model_create(m->ndm, m->ndf);
  if (!x) { G.err = xb_last_error(); return -1; }
  G.x = x;
  // 1. nodes in Domain order
  std::vector<int> tags; std::vector<double> crd;
  { NodeIter& ni = dom->getNodes(); Node* nd;
    while ((nd = ni()) != nullptr) {
      tags.push_back(nd->getTag());
      const Vector& c = nd->getCrds();
      for (int d = 0; d < m->ndm; d++) crd.push_back(c(d));
    } }
  if (xb_add_nodes(x, (int)tags.size(), tags.data(), crd.data()) < 0) { G.err = xb_last_error(); return -2; }
  // nodes created under another `model -ndf` (a quad's 2-dof nodes in a 3-dof frame model)
  { NodeIter& ni = dom->getNodes(); Node* nd;
    while ((nd = ni()) != nullptr) {
      const int k = nd->getNumberDOF(), t = nd->getTag();
      if (k > m->ndf) { G.err = "glue: a node with more dofs than the model's ndf"; return -2; }
      if (k < m->ndf && xb_set_node_ndf(x, 1, &t, k) < 0) { G.err = xb_last_error(); return -2; }
    } }
  // 2. SP constraints
  std::vector<int> spn, spd;
  { SP_ConstraintIter& si = dom->getDomainAndLoadPatternSPs(); SP_Constraint* sp;
    while ((sp = si()) != nullptr) {
      // the device path knows homogeneous constraints only (`fix`): an imposed non-zero value (`sp`) stays on the CPU
      if (sp->getValue() != 0.0) { G.err = "glue: SP_Constraint with a non-zero value: outside the device path"; return -3; }
      spn.push_back(sp->getNodeTag()); spd.push_back(sp->getDOF_Number());
    } }
  if (!spn.empty() && xb_add_sp(x, (int)spn.size(), spn.data(), spd.data()) < 0) { G.err = xb_last_error(); return -3; }
  // 2b. MP constraints: `equalDOF` only (identity constraint matrix on the same dofs -- what PlainHandler accepts)
  { MP_ConstraintIter& mi = dom->getMPs(); MP_Constraint* mp;
    while ((mp = mi()) != nullptr) {
      const ID& cd = mp->getConstrainedDOFs(); const ID& rd = mp->getRetainedDOFs(); const Matrix& C = mp->getConstraint();
      bool ident = cd.Size() == rd.Size() && C.noRows() == cd.Size() && C.noCols() == cd.Size() && !mp->isTimeVarying();
      for (int i = 0; ident && i < cd.Size(); i++) {
        if (cd(i) != rd(i)) ident = false;
        for (int j = 0; ident && j < cd.Size(); j++) if (C(i, j) != (i == j ? 1.0 : 0.0)) ident = false;
      }
      if (!ident) { G.err = "MP_Constraint that is not an equalDOF: outside the device path"; return -3; }
      std::vector<int> dofs(cd.Size());
      for (int i = 0; i < cd.Size(); i++) dofs[i] = cd(i);
      if (xb_add_equal_dof(x, mp->getNodeRetained(), mp->getNodeConstrained(), (int)dofs.size(), dofs.data()) < 0) { G.err = xb_last_error(); return -3; }
    } }
  // 3. materials and elements, one batch per (element class, material kind)
  struct Batch { std::vector<int> tag, conn, mat; std::vector<double> par; };
  std::map<std::pair<int, int>, Batch> batches;     // (xb element kind, xb material kind)
  std::map<int, int> mats_done, secs_done, unis_done;
  struct BeamKey { int sec, nip, mi; double tol; bool is3; int transf; int lobatto; };
  struct BeamRule { int tag, nip; std::vector<double> xi, wt; };
  std::vector<BeamRule> beam_rules;
  std::vector<BeamKey> beam_keys;
  auto material = [&](NDMaterial* nm, int& kind) -> int {
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int np = 0;
    if (auto* j = dynamic_cast<J2Plasticity*>(nm)) {
      kind = XB_MAT_J2PLASTICITY;
      p[0] = j->bulk; p[1] = j->shear; p[2] = j->sigma_0; p[3] = j->sigma_infty; p[4] = j->delta; p[5] = j->Hard; p[6] = j->eta; p[7] = j->rho; np = 8;
    } else if (auto* e = dynamic_cast<ElasticIsotropicMaterial*>(nm)) {
      kind = XB_MAT_ELASTIC_ISOTROPIC; p[0] = e->E; p[1] = e->v; p[2] = e->rho; np = 3;
    } else return -1;
    const int tag = nm->getTag();            // getCopy keeps the tag of the nDMaterial command
    if (!mats_done.count(tag)) {
      if (xb_add_nd_material(x, tag, kind, p, np) < 0) return -2;
      mats_done[tag] = kind;
    }
    return tag;
  };
  { ElementIter& ei = dom->getElements(); Element* el;
    while ((el = ei()) != nullptr) {
      int mk = 0;
      if (auto* b = dynamic_cast<Brick*>(el)) {
        const int mt = material(b->materialPointers[0], mk);
        if (mt < 0) { G.err = "glue: unsupported nDMaterial in a Brick"; return -4; }
        Batch& B = batches[{XB_ELE_STDBRICK, mk}];
        B.tag.push_back(el->getTag()); B.mat.push_back(mt);
        for (int a = 0; a < 8; a++) B.conn.push_back(b->connectedExternalNodes(a));
        for (int d = 0; d < 3; d++) B.par.push_back(b->b[d]);
      } else if (auto* q = dynamic_cast<FourNodeQuad*>(el)) {
        const int mt = material(q->theMaterial[0], mk);
        if (mt < 0) { G.err = "glue: unsupported nDMaterial in a FourNodeQuad"; return -4; }
        Batch& B = batches[{XB_ELE_FOURNODEQUAD, mk}];
        B.tag.push_back(el->getTag()); B.mat.push_back(mt);
        for (int a = 0; a < 4; a++) B.conn.push_back(q->connectedExternalNodes(a));
        const char* ty = q->theMaterial[0]->getType();        // the copy FourNodeQuad asked for: "PlaneStrain" | "PlaneStress"
        const bool pstress = std::strcmp(ty, "PlaneStress") == 0;
        if (!pstress && std::strcmp(ty, "PlaneStrain") != 0) { G.err = "glue: FourNodeQuad material copy is neither PlaneStrain nor PlaneStress"; return -4; }
        // FourNodeQuad takes the element's own rho instead of the material's when it is non-zero (FourNodeQuad.cpp:395-398);
        // the device path carries the material density only
        if (q->rho != 0.0) { G.err = "glue: FourNodeQuad with an element density (rho): outside the device path"; return -5; }
        const double par[6] = {q->thickness, pstress ? 1.0 : 0.0, q->pressure, q->rho, q->b[0], q->b[1]};
        B.par.insert(B.par.end(), par, par + 6);
      } else if (dynamic_cast<ForceBeamColumn2d*>(el) || dynamic_cast<ForceBeamColumn3d*>(el)) {
        // forceBeamColumn with nIP copies of one fibre section (Steel02 / Concrete02 fibres), Lobatto integration,
        // geomTransf Linear without joint offsets
        auto* b2 = dynamic_cast<ForceBeamColumn2d*>(el); auto* b3 = dynamic_cast<ForceBeamColumn3d*>(el);
        const int nsec = b2 ? b2->numSections : b3->numSections;
        SectionForceDeformation** secs = b2 ? b2->sections : b3->sections;
        BeamIntegration* bi = b2 ? b2->beamIntegr : b3->beamIntegr;
        CrdTransf* ct = b2 ? b2->crdTransf : b3->crdTransf;
        // any BeamIntegration whose sections are all the same: the locations and weights are the reference object's own
        if (!dynamic_cast<LobattoBeamIntegration*>(bi)) {
          BeamRule br; br.tag = el->getTag(); br.nip = nsec; br.xi.resize(nsec); br.wt.resize(nsec);
          const double Lr = ct->getInitialLength();
          bi->getSectionLocations(nsec, Lr, br.xi.data());
          bi->getSectionWeights(nsec, Lr, br.wt.data());
          beam_rules.push_back(br);
        }
        // geomTransf Linear or PDelta, without joint offsets
        int transf = -1;
        double off2[4] = {0.0, 0.0, 0.0, 0.0};      // 2D: -jntOffset goes along
        if (auto* t = dynamic_cast<LinearCrdTransf2d*>(ct)) {
          transf = 0;
          if (t->nodeIOffset) { off2[0] = t->nodeIOffset[0]; off2[1] = t->nodeIOffset[1]; }
          if (t->nodeJOffset) { off2[2] = t->nodeJOffset[0]; off2[3] = t->nodeJOffset[1]; }
        }
        double off3[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (auto* t = dynamic_cast<LinearCrdTransf3d*>(ct)) {
          transf = 0;
          if (t->nodeIOffset) for (int q = 0; q < 3; q++) off3[q] = t->nodeIOffset[q];
          if (t->nodeJOffset) for (int q = 0; q < 3; q++) off3[3 + q] = t->nodeJOffset[q];
        }
        else if (auto* t = dynamic_cast<PDeltaCrdTransf2d*>(ct)) {
          transf = 1;
          if (t->nodeIOffset) { off2[0] = t->nodeIOffset[0]; off2[1] = t->nodeIOffset[1]; }
          if (t->nodeJOffset) { off2[2] = t->nodeJOffset[0]; off2[3] = t->nodeJOffset[1]; }
        }
        else if (auto* t = dynamic_cast<CorotCrdTransf2d*>(ct)) {
          if (t->nodeOffsets) { G.err = "glue: geomTransf Corotational with joint offsets: outside the device path"; return -5; }
          transf = 2;
        }
        else if (auto* t = dynamic_cast<PDeltaCrdTransf3d*>(ct)) {
          transf = 1;
          if (t->nodeIOffset) for (int q = 0; q < 3; q++) off3[q] = t->nodeIOffset[q];
          if (t->nodeJOffset) for (int q = 0; q < 3; q++) off3[3 + q] = t->nodeJOffset[q];
        }
        if (transf < 0) { G.err = "glue: geomTransf other than Linear / PDelta / Corotational (2D): outside the device path"; return -5; }
        for (int i = 1; i < nsec; i++) if (secs[i]->getTag() != secs[0]->getTag()) { G.err = "glue: sections of one element differ"; return -5; }
        const int stag = secs[0]->getTag();
        if (!secs_done.count(stag)) {
          auto uniaxial = [&](UniaxialMaterial* um) -> int {
            const int t = um->getTag();
            if (unis_done.count(t)) return t;
            double p[12] = {0}; int kind, np;
            if (auto* s2 = dynamic_cast<Steel02*>(um)) {
              kind = XB_UNI_STEEL02; np = 11;
              const double q[11] = {s2->Fy, s2->E0, s2->b, s2->R0, s2->cR1, s2->cR2, s2->a1, s2->a2, s2->a3, s2->a4, s2->sigini};
              std::memcpy(p, q, sizeof q);
            } else if (auto* c2 = dynamic_cast<Concrete02*>(um)) {
              kind = XB_UNI_CONCRETE02; np = 7;
              const double q[7] = {c2->fc, c2->epsc0, c2->fcu, c2->epscu, c2->rat, c2->ft, c2->Ets};
              std::memcpy(p, q, sizeof q);
            } else if (auto* s1 = dynamic_cast<Steel01*>(um)) {
              kind = XB_UNI_STEEL01; np = 7;
              const double q[7] = {s1->fy, s1->E0, s1->b, s1->a1, s1->a2, s1->a3, s1->a4};
              std::memcpy(p, q, sizeof q);
            } else if (auto* c1 = dynamic_cast<Concrete01*>(um)) {
              kind = XB_UNI_CONCRETE01; np = 4;
              const double q[4] = {c1->fpc, c1->epsc0, c1->fpcu, c1->epscu};
              std::memcpy(p, q, sizeof q);
            } else if (auto* pp = dynamic_cast<ElasticPPMaterial*>(um)) {   // (the yield strains back out of the yield stresses, as getCopy does)
              if (pp->ep != 0.0) return -1;
              kind = XB_UNI_ELASTICPP; np = 4;
              const double q[4] = {pp->E, pp->fyp / pp->E, pp->fyn / pp->E, pp->ezero};
              std::memcpy(p, q, sizeof q);
            } else if (auto* em = dynamic_cast<ElasticMaterial*>(um)) {
              kind = XB_UNI_ELASTIC; np = 3;
              const double q[3] = {em->Epos, em->eta, em->Eneg};
              std::memcpy(p, q, sizeof q);
            } else return -1;
            if (xb_add_uniaxial_material(x, t, kind, p, np) < 0) return -1;
            unis_done[t] = 1;
            return t;
          };
          std::vector<double> y, z, A; std::vector<int> mt;
          if (auto* f2 = dynamic_cast<FiberSection2d*>(secs[0])) {
            for (int f = 0; f < f2->numFibers; f++) {
              y.push_back(f2->matData[2 * f]); A.push_back(f2->matData[2 * f + 1]);
              const int t = uniaxial(f2->theMaterials[f]); if (t < 0) { G.err = "glue: unsupported fibre material"; return -5; }
              mt.push_back(t);
            }
            if (xb_add_fiber_section(x, stag, (int)y.size(), y.data(), A.data(), mt.data()) < 0) { G.err = xb_last_error(); return -6; }
          } else if (auto* f3 = dynamic_cast<FiberSection3d*>(secs[0])) {
            for (int f = 0; f < f3->numFibers; f++) {
              y.push_back(f3->matData[3 * f]); z.push_back(f3->matData[3 * f + 1]); A.push_back(f3->matData[3 * f + 2]);
              const int t = uniaxial(f3->theMaterials[f]); if (t < 0) { G.err = "glue: unsupported fibre material"; return -5; }
              mt.push_back(t);
            }
            auto* tor = dynamic_cast<ElasticMaterial*>(f3->theTorsion);
            if (!tor) { G.err = "glue: torsion other than an ElasticMaterial"; return -5; }
            if (xb_add_fiber_section3d(x, stag, (int)y.size(), y.data(), z.data(), A.data(), mt.data(), tor->getInitialTangent()) < 0) { G.err = xb_last_error(); return -6; }
          } else if (auto* ag = dynamic_cast<SectionAggregator*>(secs[0])) {
            // section Aggregator of uniaxial materials only (configs[0]: Elastic on P, Steel01 on Mz)
            if (ag->theSection) { G.err = "glue: section Aggregator on top of another section"; return -5; }
            std::vector<int> codes;
            for (int f = 0; f < ag->numMats; f++) {
              const int t = uniaxial(ag->theAdditions[f]); if (t < 0) { G.err = "glue: unsupported material in a section Aggregator"; return -5; }
              mt.push_back(t); codes.push_back((*ag->matCodes)(f));
            }
            if (xb_add_section_aggregator(x, stag, (int)mt.size(), mt.data(), codes.data()) < 0) { G.err = xb_last_error(); return -6; }
          } else { G.err = "glue: section other than a fibre section or an Aggregator of uniaxial materials"; return -5; }
          secs_done[stag] = 1;
        }
        // one batch per (section, nIP, maxIters, tol): key them through the map's second index
        const int maxIters = b2 ? b2->maxIters : b3->maxIters; const double tol = b2 ? b2->tol : b3->tol;
        int key = -1;
        const int lob = dynamic_cast<LobattoBeamIntegration*>(bi) ? 1 : 0;      // (a batch integrates by Lobatto or by per-element rules)
        for (size_t q = 0; q < beam_keys.size(); q++)
          if (beam_keys[q].sec == stag && beam_keys[q].nip == nsec && beam_keys[q].mi == maxIters && beam_keys[q].tol == tol && beam_keys[q].is3 == (b3 != nullptr) && beam_keys[q].transf == transf && beam_keys[q].lobatto == lob) key = (int)q;
        if (key < 0) { beam_keys.push_back({stag, nsec, maxIters, tol, b3 != nullptr, transf, lob}); key = (int)beam_keys.size() - 1; }
        Batch& B = batches[{b3 ? XB_ELE_FORCEBEAMCOLUMN3D : XB_ELE_FORCEBEAMCOLUMN2D, 1000 + key}];
        B.tag.push_back(el->getTag()); B.mat.push_back(stag);
        const ID& en = el->getExternalNodes();
        B.conn.push_back(en(0)); B.conn.push_back(en(1));
        B.par.push_back(nsec); B.par.push_back(maxIters); B.par.push_back(tol);
        if (b3) {   // the local z axis serves as vecxz: y = z ^ x, so the same triad comes out
          Vector xa(3), ya(3), za(3);
          ct->getLocalAxes(xa, ya, za);
          for (int d = 0; d < 3; d++) B.par.push_back(za(d));
        }
        B.par.push_back(transf);
        B.par.push_back(b2 ? b2->rho : b3->rho);      // -mass: lumped, travels as nodal mass on the device
        if (b2) for (int q = 0; q < 4; q++) B.par.push_back(off2[q]);
        else for (int q = 0; q < 6; q++) B.par.push_back(off3[q]);
      } else { G.err = "glue: element class outside the device path (keep the CPU integrator)"; return -5; }
    } }
  for (auto& kv : batches) {
    Batch& B = kv.second;
    const int ek = kv.first.first;
    const int stride = ek == XB_ELE_STDBRICK ? 3 : (ek == XB_ELE_FORCEBEAMCOLUMN2D ? 9 : (ek == XB_ELE_FORCEBEAMCOLUMN3D ? 14 : 6));
    if (xb_add_elements(x, kv.first.first, (int)B.tag.size(), B.tag.data(), B.conn.data(), B.mat.data(), B.par.data(), stride) < 0) {
      G.err = xb_last_error(); return -6;
    }
  }
  for (const BeamRule& br : beam_rules)
    if (xb_set_beam_integration(x, 1, &br.tag, br.nip, br.xi.data(), br.wt.data()) < 0) { G.err = xb_last_error(); return -6; }
  // 4. nodal loads of the load patterns.  The device applies lambda(t) * P with lambda = the domain time, i.e. every
  // pattern must carry a Linear series with factor 1, must not have been frozen by loadConst, and must hold nodal loads
  // only -- anything else is refused here rather than silently dropped (element loads, other series: keep the CPU integrator)
  { LoadPatternIter& pi = dom->getLoadPatterns(); LoadPattern* lp;
    while ((lp = pi()) != nullptr) {
      LinearSeries* ls = dynamic_cast<LinearSeries*>(lp->theSeries);
      if (!ls || ls->cFactor != 1.0) { G.err = "glue: load pattern whose TimeSeries is not Linear with factor 1: outside the device path"; return -7; }
      if (lp->isConstant) { G.err = "glue: load pattern frozen by loadConst before the first analysis: outside the device path"; return -7; }
      G.pattern_state[lp->getTag()] = 0;
      { ElementalLoadIter& eli = lp->getElementalLoads(); ElementalLoad* el;
        while ((el = eli()) != nullptr) {   // `eleLoad -beamUniform` on the force beams goes along; every other element load is refused
          int type; const Vector& data = el->getData(type, 1.0);
          const int et = el->getElementTag();
          Element* ele = dom->getElement(et);
          double w[3] = {0.0, 0.0, 0.0};
          if (type == LOAD_TAG_Beam2dUniformLoad && dynamic_cast<ForceBeamColumn2d*>(ele)) { w[0] = data(0); w[2] = data(1); }
          else if (type == LOAD_TAG_Beam3dUniformLoad && dynamic_cast<ForceBeamColumn3d*>(ele)) { w[0] = data(0); w[1] = data(1); w[2] = data(2); }
          else if (type == LOAD_TAG_Beam2dPointLoad && dynamic_cast<ForceBeamColumn2d*>(ele)) {   // Ptrans, Paxial, x/L
            const double p4[4] = {data(0), 0.0, data(1), data(2)};
            if (xb_add_beam_point_loads(x, 1, &et, p4) < 0) { G.err = xb_last_error(); return -7; }
            continue;
          } else if (type == LOAD_TAG_Beam3dPointLoad && dynamic_cast<ForceBeamColumn3d*>(ele)) {   // Py, Pz, Px, x/L
            const double p4[4] = {data(0), data(1), data(2), data(3)};
            if (xb_add_beam_point_loads(x, 1, &et, p4) < 0) { G.err = xb_last_error(); return -7; }
            continue;
          }
          else if (type == LOAD_TAG_Beam2dPartialUniformLoad && dynamic_cast<ForceBeamColumn2d*>(ele)) {   // wTa, wTb, wAa, wAb, a/L, b/L
            const double p8[8] = {data(0), data(1), data(2), data(3), data(4), data(5), 0.0, 0.0};
            if (xb_add_beam_partial_loads(x, 1, &et, p8) < 0) { G.err = xb_last_error(); return -7; }
            continue;
          }
          else if (type == LOAD_TAG_Beam3dPartialUniformLoad && dynamic_cast<ForceBeamColumn3d*>(ele)) {   // wya, wza, waa, a/L, b/L, wyb, wzb, wab
            const double p8[8] = {data(0), data(5), data(2), data(7), data(3), data(4), data(1), data(6)};
            if (xb_add_beam_partial_loads(x, 1, &et, p8) < 0) { G.err = xb_last_error(); return -7; }
            continue;
          }
          else { G.err = "glue: ElementalLoad other than -beamUniform / -beamPoint on a forceBeamColumn: outside the device path"; return -7; }
          if (xb_add_beam_uniform_loads(x, 1, &et, w) < 0) { G.err = xb_last_error(); return -7; }
        } }
      NodalLoadIter& li = lp->getNodalLoads(); NodalLoad* nl;
      while ((nl = li()) != nullptr) {
        int type; const Vector& v = nl->getData(type);
        std::vector<double> vals(m->ndf, 0.0);
        for (int d = 0; d < m->ndf && d < v.Size(); d++) vals[d] = v(d);
        const int nt = nl->getNodeTag();
        if (xb_add_nodal_loads(x, 1, &nt, vals.data()) < 0) { G.err = xb_last_error(); return -7; }
      }
    } }
  // 4b. `mass` command (Node::getMass, diagonal) and `rayleigh` (Element / Node factors)
  { std::vector<int> mt; std::vector<double> mv;
    NodeIter& ni = dom->getNodes(); Node* nd;
    while ((nd = ni()) != nullptr) {
      const Matrix& M = nd->getMass();
      bool any = false;
      for (int d = 0; d < m->ndf && d < M.noRows(); d++) if (M(d, d) != 0.0) any = true;
      if (!any) continue;
      mt.push_back(nd->getTag());
      for (int d = 0; d < m->ndf; d++) mv.push_back(d < M.noRows() ? M(d, d) : 0.0);
    }
    if (!mt.empty() && xb_set_nodal_mass(x, (int)mt.size(), mt.data(), mv.data()) < 0) { G.err = xb_last_error(); return -7; } }
  { ElementIter& ei = dom->getElements(); Element* el = ei();
    const double rf[4] = {el ? el->alphaM : 0.0, el ? el->betaK : 0.0, el ? el->betaK0 : 0.0, el ? el->betaKc : 0.0};
    if (el && (rf[0] != 0.0 || rf[1] != 0.0 || rf[2] != 0.0 || rf[3] != 0.0))
      if (xb_set_rayleigh(x, rf[0], rf[1], rf[2], rf[3]) < 0) { G.err = xb_last_error(); return -7; }
    // one set of factors for the whole model (the `rayleigh` command): per-element factors (region -rayleigh) are refused
    bool same = true;
    while (el) { if (el->alphaM != rf[0] || el->betaK != rf[1] || el->betaK0 != rf[2] || el->betaKc != rf[3]) same = false; el = ei(); }
    if (!same) { G.err = "glue: Rayleigh factors differ between elements: outside the device path"; return -7; } }
  // 5. the same numberer / SOE as the analysis; the numbering must be the reference's own
  const int neq = xb_setup(x, numberer, soeKind);
  if (neq < 0) { G.err = xb_last_error(); return -8; }
  if (neq != m->soe->getNumEqn()) { G.err = "glue: equation count differs from the reference's"; return -9; }
  if (m->bsoe) {   // system BandGeneral: the band widths and the array length must be the reference's
    int kl = -1, ku = -1;
    xb_get_band(x, &kl, &ku);
    if (kl != m->bsoe->sub() || ku != m->bsoe->super() || xb_a_size(x) != m->bsoe->asize()) { G.err = "glue: band widths differ from the reference BandGenLinSOE's"; return -9; }
  } else {   // xb_form_tangent writes xb_nnz doubles straight into the SOE's A: the two patterns must be the same
    const long long nz = xb_nnz(x);
    if (nz != (long long)m->ptr()[neq]) { G.err = "glue: non-zero count differs from the reference SOE's"; return -9; }
    std::vector<long long> xp((size_t)neq + 1); std::vector<int> xi((size_t)nz);
    xb_get_pattern(x, xp.data(), xi.data());
    for (int i = 0; i <= neq; i++) if (xp[i] != (long long)m->ptr()[i]) { G.err = "glue: sparse pattern (pointers) differs from the reference SOE's"; return -9; }
    for (long long i = 0; i < nz; i++) if (xi[i] != m->idx()[i]) { G.err = "glue: sparse pattern (indices) differs from the reference SOE's"; return -9; }
  }
  std::vector<int> xt(tags.size()), ids(tags.size() * m->ndf);
  xb_get_node_tags(x, xt.data()); xb_get_ids(x, ids.data());
  // `constraints Transformation`: with homogeneous SPs and identity MP constraints the TransformationConstraintHandler numbers
  // the same equations and gives the same pattern as PlainHandler (tests/test_oracle.py); only the layout of a constrained
  // node's ID differs -- TransformationDOF_Group::getID is [the node's own unconstrained dofs..., the retained node's
  // retained dofs...] (TransformationDOF_Group.cpp:60-110, doneID :921-945).  Its enforceSPs() also calls Element::update()
  // once more on every element next to a constrained node at each applyLoad (TransformationConstraintHandler.cpp:462-483):
  // after a commit that leaves a yielded J2 point with its elastic tangent for the first iteration of the next step, so the
  // reference's Newton histories differ between its two handlers.  The device does the same second update when told
  // (`constraints_transformation`); a force-based beam iterates again from where it stood, there as here.
  std::map<int, size_t> row_of;
  for (size_t i = 0; i < xt.size(); i++) row_of[xt[i]] = i;
  const bool transf_handler = dynamic_cast<TransformationConstraintHandler*>(m->handler) != nullptr;
  if (transf_handler) {
    if (xb_set_option(x, "constraints_transformation", 1) < 0) { G.err = xb_last_error(); return -10; }
  }
  for (size_t i = 0; i < xt.size(); i++) {
    DOF_Group* grp = dom->getNode(xt[i])->getDOF_GroupPtr();
    const ID& rid = grp->getID();
    TransformationDOF_Group* tg = transf_handler ? dynamic_cast<TransformationDOF_Group*>(grp) : nullptr;
    if (tg && tg->theMP) {
      const ID& cd = tg->theMP->getConstrainedDOFs(); const ID& rd = tg->theMP->getRetainedDOFs();
      std::vector<int> want;
      const int nnd = dom->getNode(xt[i])->getNumberDOF();
      for (int d = 0; d < nnd; d++) if (cd.getLocation(d) < 0) want.push_back(ids[i * m->ndf + d]);
      auto it = row_of.find(tg->theMP->getNodeRetained());
      if (it == row_of.end()) { G.err = "glue: retained node of an MP_Constraint not in the model"; return -10; }
      for (int j = 0; j < rd.Size(); j++) want.push_back(ids[it->second * m->ndf + rd(j)]);
      if ((int)want.size() != rid.Size()) { G.err = "glue: DOF numbering differs from the reference's (constrained node)"; return -10; }
      for (int k = 0; k < rid.Size(); k++) if (rid(k) != want[k]) { G.err = "glue: DOF numbering differs from the reference's (constrained node)"; return -10; }
      continue;
    }
    for (int d = 0; d < m->ndf; d++)
      if ((d < rid.Size() ? rid(d) : -1) != ids[i * m->ndf + d]) { G.err = "glue: DOF numbering differs from the reference's"; return -10; }
  }
  if (xb_device_init(x, device, nullptr) < 0) { G.err = xb_last_error(); return -11; }
  return neq;
}

// INTEGRATION.md "B200LoadControl"
class B200LoadControl : public LoadControl {
 public:
  B200LoadControl(double dl) : LoadControl(dl, 1, dl, dl) {}
  xb_model* x = nullptr;
  RefModel* rm = nullptr;
  long calls[4] = {0, 0, 0, 0};
  double* soeA() { return rm->bsoe ? rm->bsoe->a() : (rm->rsoe ? rm->rsoe->A : rm->csoe->A); }
  Vector& soeB() { return rm->bsoe ? rm->bsoe->bvec() : (rm->rsoe ? rm->rsoe->B : rm->csoe->B); }

  int formTangent(int statFlag) override {
    if (statFlag != CURRENT_TANGENT) return LoadControl::formTangent(statFlag);
    statusFlag = statFlag;
    this->getLinearSOE()->zeroA();              // resets the solver's "factored" flag; every entry is overwritten below
    calls[0]++;
    return xb_form_tangent(x, soeA());
  }
  int formTangent(int statFlag, double iFactor, double cFactor) override {
    if (statFlag != CURRENT_TANGENT) return LoadControl::formTangent(statFlag, iFactor, cFactor);
    return this->formTangent(statFlag);
  }
  // LoadControl::newStep (final) ends in AnalysisModel::applyLoadDomain -- under `constraints Transformation` that is where
  // the handler updates the constrained elements once more.  The first formUnbalance of a step follows it directly
  // (NewtonRaphson::solveCurrentStep), so that call carries xb_apply_load; the later ones only re-read the domain time
  bool step_begins = true;
  // StaticIntegrator::formUnbalance is final: zeroB(); formElementResidual(); formNodalUnbalance()
  int formElementResidual() override {
    calls[1]++;
    const double t = this->getAnalysisModel()->getCurrentDomainTime();
    if ((step_begins ? xb_apply_load(x, t) : xb_set_load_factor(x, t)) < 0) return -1;
    step_begins = false;
    Vector& B = soeB();
    return xb_form_unbalance(x, &B(0));         // elements and nodal loads, the whole right-hand side
  }
  int formNodalUnbalance() override { return 0; }
  int update(const Vector& dU) override {
    AnalysisModel* am = this->getAnalysisModel();
    LinearSOE* soe = this->getLinearSOE();
    am->incrDisp(dU);                            // the reference's nodes follow (recorders read them); no Element::update
    calls[2]++;
    std::vector<double> du(dU.Size());
    for (int i = 0; i < dU.Size(); i++) du[i] = dU(i);
    if (xb_incr_trial_disp(x, du.data()) < 0 || xb_update(x) < 0) return -1;
    soe->setX(dU);
    numIncrLastStep++;
    return 0;
  }
  int commit() override {
    calls[3]++;
    step_begins = true;
    if (xb_commit(x) < 0) return -1;
    return LoadControl::commit();
  }
  // a failed step: BasicAnalysisBuilder::analyzeStatic / analyzeTransient call Domain::revertToLastCommit and then this
  // (BasicAnalysisBuilder.cpp:372-413, 488-519); the device state goes back with it
  int revertToLastStep() override {
    step_begins = true;
    const int rc = LoadControl::revertToLastStep();
    if (x && xb_revert_to_last_commit(x) < 0) return -1;
    return rc;
  }
};

// `integrator DisplacementControl`: newStep / update as DisplacementControl.cpp:121,210 (no sensitivities), with
// AnalysisModel::updateDomain replaced; the two solves per iteration stay in the reference's SOE
class B200DisplacementControl : public DisplacementControl {
 public:
  B200DisplacementControl(int node, int dof, double incr, Domain* d) : DisplacementControl(node, dof, incr, d, 1, incr, incr) {}
  xb_model* x = nullptr;
  RefModel* rm = nullptr;
  long calls[4] = {0, 0, 0, 0};
  double* soeA() { return rm->bsoe ? rm->bsoe->a() : (rm->rsoe ? rm->rsoe->A : rm->csoe->A); }
  Vector& soeB() { return rm->bsoe ? rm->bsoe->bvec() : (rm->rsoe ? rm->rsoe->B : rm->csoe->B); }
  int push(const Vector& dU, double lambda) {          // incrDisp + applyLoadDomain + updateDomain
    AnalysisModel* am = this->getAnalysisModel();
    am->incrDisp(dU);
    am->applyLoadDomain(lambda);
    std::vector<double> du(dU.Size());
    for (int i = 0; i < dU.Size(); i++) du[i] = dU(i);
    if (xb_incr_trial_disp(x, du.data()) < 0 || xb_apply_load(x, lambda) < 0 || xb_update(x) < 0) return -1;
    return 0;
  }
  int newStep() override {
    if (theDofID == -1) return -1;
    AnalysisModel* theModel = this->getAnalysisModel();
    LinearSOE* theLinSOE = this->getLinearSOE();
    double factor = pow(specNumIncrStep / numIncrLastStep, 1.0);
    theIncrement *= factor;
    if (theIncrement < minIncrement) theIncrement = minIncrement;
    else if (theIncrement > maxIncrement) theIncrement = maxIncrement;
    currentLambda = theModel->getCurrentDomainTime();
    this->formTangent(tangFlag);
    theLinSOE->setB(*phat);
    if (theLinSOE->solve() < 0) return -1;
    (*deltaUhat) = theLinSOE->getX();
    const double dUahat = (*deltaUhat)(theDofID);
    if (dUahat == 0.0) return -1;
    const double dlambda = theIncrement / dUahat;
    deltaLambdaStep = dlambda;
    currentLambda += dlambda;
    (*deltaU) = *deltaUhat;
    (*deltaU) *= dlambda;
    (*deltaUstep) = (*deltaU);
    if (push(*deltaU, currentLambda) < 0) return -1;
    numIncrLastStep = 0;
    return 0;
  }
  int update(const Vector& dU) override {
    if (theDofID == -1) return -1;
    LinearSOE* theLinSOE = this->getLinearSOE();
    (*deltaUbar) = dU;
    const double dUabar = (*deltaUbar)(theDofID);
    theLinSOE->setB(*phat);
    theLinSOE->solve();
    (*deltaUhat) = theLinSOE->getX();
    const double dUahat = (*deltaUhat)(theDofID);
    if (dUahat == 0.0) return -1;
    dLambda = -dUabar / dUahat;
    (*deltaU) = (*deltaUbar);
    deltaU->addVector(1.0, *deltaUhat, dLambda);
    (*deltaUstep) += *deltaU;
    deltaLambdaStep += dLambda;
    currentLambda += dLambda;
    calls[2]++;
    if (push(*deltaU, currentLambda) < 0) return -1;
    theLinSOE->setX(*deltaU);
    numIncrLastStep++;
    return 0;
  }
  int formTangent(int statFlag) override {
    if (!x || statFlag != CURRENT_TANGENT) return DisplacementControl::formTangent(statFlag);
    statusFlag = statFlag;
    this->getLinearSOE()->zeroA();
    calls[0]++;
    return xb_form_tangent(x, soeA());
  }
  int formTangent(int statFlag, double iFactor, double cFactor) override {
    if (!x || statFlag != CURRENT_TANGENT) return DisplacementControl::formTangent(statFlag, iFactor, cFactor);
    return this->formTangent(statFlag);
  }
  int formElementResidual() override {
    if (!x) return DisplacementControl::formElementResidual();   // domainChanged before the device model exists
    calls[1]++;
    if (xb_set_load_factor(x, this->getAnalysisModel()->getCurrentDomainTime()) < 0) return -1;
    Vector& B = soeB();
    return xb_form_unbalance(x, &B(0));
  }
  int formNodalUnbalance() override { return x ? 0 : DisplacementControl::formNodalUnbalance(); }
  int commit() override {
    calls[3]++;
    if (xb_commit(x) < 0) return -1;
    return DisplacementControl::commit();
  }
  // a failed step: BasicAnalysisBuilder::analyzeStatic / analyzeTransient call Domain::revertToLastCommit and then this
  // (BasicAnalysisBuilder.cpp:372-413, 488-519); the device state goes back with it
  int revertToLastStep() override {
    const int rc = DisplacementControl::revertToLastStep();
    if (x && xb_revert_to_last_commit(x) < 0) return -1;
    return rc;
  }
};

// the transient counterpart: Newmark (displacement unknown).  newStep / update keep the reference's own U, Udot,
// Udotdot vectors and its nodes up to date (AnalysisModel::setVel / setAccel / setResponse are node-level), and
// replace AnalysisModel::updateDomain -> Domain::update, formTangent and formUnbalance.
class B200Newmark : public Newmark {
 public:
  B200Newmark(double g, double b) : Newmark(g, b) {}
  xb_model* x = nullptr;
  RefModel* rm = nullptr;
  long calls[4] = {0, 0, 0, 0};
  double* soeA() { return rm->bsoe ? rm->bsoe->a() : (rm->rsoe ? rm->rsoe->A : rm->csoe->A); }
  Vector& soeB() { return rm->bsoe ? rm->bsoe->bvec() : (rm->rsoe ? rm->rsoe->B : rm->csoe->B); }

  int newStep(double deltaT) override {                 // Newmark::newStep (Newmark.cpp:105), unknown = Displacement
    if (deltaT <= 0.0 || beta == 0 || U == nullptr || unknown != 1) return -1;
    AnalysisModel* theModel = this->getAnalysisModel();
    c1 = 1.0; c2 = gamma / (beta * deltaT); c3 = 1.0 / (beta * deltaT * deltaT);
    (*Ut) = *U; (*Utdot) = *Udot; (*Utdotdot) = *Udotdot;
    const double a1 = (1.0 - gamma / beta), a2 = deltaT * (1.0 - 0.5 * gamma / beta);
    Udot->addVector(a1, *Utdotdot, a2);
    const double a3 = -1.0 / (beta * deltaT), a4 = 1.0 - 0.5 / beta;
    Udotdot->addVector(a4, *Utdot, a3);
    theModel->setVel(*Udot);
    theModel->setAccel(*Udotdot);
    double time = theModel->getCurrentDomainTime();
    time += deltaT;
    theModel->applyLoadDomain(time);                    // in place of updateDomain(time, dT): loads on the reference's nodes ...
    if (xb_set_transient_factors(x, c1, c2, c3) < 0 || xb_newmark_predict(x, a1, a2, a3, a4) < 0 ||
        xb_apply_load(x, time) < 0 || xb_update(x) < 0) return -2;   // ... and Domain::update on the device
    return 0;
  }
  int update(const Vector& deltaU) override {           // Newmark::update (Newmark.cpp:411)
    AnalysisModel* theModel = this->getAnalysisModel();
    (*U) += deltaU;
    Udot->addVector(1.0, deltaU, c2);
    Udotdot->addVector(1.0, deltaU, c3);
    theModel->setResponse(*U, *Udot, *Udotdot);
    calls[2]++;
    std::vector<double> du(deltaU.Size());
    for (int i = 0; i < deltaU.Size(); i++) du[i] = deltaU(i);
    if (xb_incr_trial_response(x, du.data(), 1.0, c2, c3) < 0 || xb_update(x) < 0) return -1;
    return 0;
  }
  int formTangent(int statFlag) override {              // TransientIntegrator::formTangent (TransientIntegrator.cpp:61)
    if (statFlag != CURRENT_TANGENT) return Newmark::formTangent(statFlag);
    statusFlag = statFlag;
    this->getLinearSOE()->zeroA();
    calls[0]++;
    return xb_form_tangent(x, soeA());
  }
  int formUnbalance() override {                        // TransientIntegrator::formUnbalance (:114)
    calls[1]++;
    this->getLinearSOE()->zeroB();
    Vector& B = soeB();
    return xb_form_unbalance(x, &B(0));
  }
  int commit() override {
    calls[3]++;
    if (xb_commit(x) < 0) return -1;
    return Newmark::commit();
  }
  // a failed step: BasicAnalysisBuilder::analyzeStatic / analyzeTransient call Domain::revertToLastCommit and then this
  // (BasicAnalysisBuilder.cpp:372-413, 488-519); the device state goes back with it
  int revertToLastStep() override {
    const int rc = Newmark::revertToLastStep();
    if (x && xb_revert_to_last_commit(x) < 0) return -1;
    return rc;
  }
};

}  // namespace

extern "C" {

// analysis set-up as ref_setup_dispcontrol, with the device-backed DisplacementControl; returns numEqn
int glue_setup_dispcontrol(void* h, int numberer, int soeKind, int node, int dof, double incr, int testKind, double tol, int maxIter, int device) {
  RefModel* m = (RefModel*)h;
  B200DisplacementControl* li = new B200DisplacementControl(node, dof, incr, m->domain);
  m->sinteg = li; m->integ = li;
  const int neq = ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
  if (neq < 0) return neq;
  Glue& G = g_glue[h];
  const int rc = domain_to_xb(m, numberer, soeKind, device, G);
  if (rc < 0) { fprintf(stderr, "glue: %s\n", G.err.c_str()); return -100 + rc; }
  li->x = G.x; li->rm = m;
  if (li->domainChanged() < 0) return -200;      // the reference load vector phat, now through the device path
  return neq;
}
// analysis set-up as ref_setup_transient, with the device-backed Newmark; returns numEqn
int glue_setup_newmark(void* h, int numberer, int soeKind, double gamma, double beta, int testKind, double tol, int maxIter, int device) {
  RefModel* m = (RefModel*)h;
  B200Newmark* ni = new B200Newmark(gamma, beta);
  m->tinteg = ni; m->integ = ni;
  const int neq = ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
  if (neq < 0) return neq;
  Glue& G = g_glue[h];
  const int rc = domain_to_xb(m, numberer, soeKind, device, G);
  if (rc < 0) { fprintf(stderr, "glue: %s\n", G.err.c_str()); return -100 + rc; }
  ni->x = G.x; ni->rm = m;
  return neq;
}

// analysis set-up as ref_setup, with the device-backed LoadControl; returns numEqn
int glue_setup_loadcontrol(void* h, int numberer, int soeKind, double dlambda, int testKind, double tol, int maxIter, int device) {
  RefModel* m = (RefModel*)h;
  B200LoadControl* li = new B200LoadControl(dlambda);
  m->sinteg = li; m->integ = li;
  const int neq = ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
  if (neq < 0) return neq;
  Glue& G = g_glue[h];
  const int rc = domain_to_xb(m, numberer, soeKind, device, G);
  if (rc < 0) { fprintf(stderr, "glue: %s\n", G.err.c_str()); return -100 + rc; }
  li->x = G.x; li->rm = m;
  return neq;
}
const char* glue_last_error(void* h) { return g_glue[h].err.c_str(); }
// frees the device model of a glued reference model (call before ref_destroy)
void glue_destroy(void* h) {
  auto it = g_glue.find(h);
  if (it == g_glue.end()) return;
  if (it->second.x) xb_model_destroy(it->second.x);
  g_glue.erase(it);
}
// how often the reference's algorithm went through each replaced loop: formTangent, formUnbalance, update, commit
void glue_call_counts(void* h, long* out) {
  B200LoadControl* li = dynamic_cast<B200LoadControl*>(((RefModel*)h)->sinteg);
  B200Newmark* ni = dynamic_cast<B200Newmark*>(((RefModel*)h)->tinteg);
  B200DisplacementControl* di = dynamic_cast<B200DisplacementControl*>(((RefModel*)h)->sinteg);
  for (int i = 0; i < 4; i++) out[i] = li ? li->calls[i] : (ni ? ni->calls[i] : (di ? di->calls[i] : -1));
}
long long glue_launch_count(void* h) { return xb_launch_count(g_glue[h].x); }
// trial displacements of the device model, [nn][ndf] in Domain order
int glue_get_trial_disp(void* h, double* u) { return xb_get_trial_disp(g_glue[h].x, u); }

}  // extern "C"
