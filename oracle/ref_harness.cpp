// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product path.
//
// C-callable driver around the UNMODIFIED reference classes, compiled from the
// sources under /root/reference by oracle/ref_build.mk into
// oracle/_ref/libxara_ref.a.  This file is ours; it only *calls* the
// reference's public C++ API (Domain, Node, Brick, FourNodeQuad, NDMaterial,
// PlainHandler, PlainNumberer / DOF_Numberer+RCM, SparseGenRowLinSOE,
// LoadControl, NewtonRaphson ...).  It plays the part that
// runtime/runtime/BasicAnalysisBuilder.cpp plays in the reference (that file
// needs Tcl and is not built): domainChanged() (BasicAnalysisBuilder.cpp:225)
// and analyzeStatic() (BasicAnalysisBuilder.cpp:337) are restated below in
// their Increment / Iterate / Commit order.
//
// Used by tests/ to pin the oracle (oracle/xara_oracle.c) and by
// `bench.py --impl reference` / cpu_baseline as the "reference" CPU arm.

#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <vector>
#include <array>

#include <Domain.h>
#include <Node.h>
#include <NodeIter.h>
#include <Element.h>
#include <ElementIter.h>
#include <Vector.h>
#include <Matrix.h>
#include <ID.h>
#include <SP_Constraint.h>
#include <MP_Constraint.h>
#include <LoadPattern.h>
#include <NodalLoad.h>
#include <LinearSeries.h>

#include <NDMaterial.h>
#include <ElasticIsotropicMaterial.h>
#include <J2Plasticity.h>
#include <UniaxialMaterial.h>

#include <Brick.h>
#include <FourNodeQuad.h>
#include <Steel02.h>
#include <Concrete02.h>
#include <FiberSection2d.h>
#include <SectionAggregator.h>
#include <PDeltaCrdTransf2d.h>
#include <CorotCrdTransf2d.h>
#include <LegendreBeamIntegration.h>
#include <RadauBeamIntegration.h>
#include <NewtonCotesBeamIntegration.h>
#include <TrapezoidalBeamIntegration.h>
#include <PDeltaCrdTransf3d.h>
#include <Steel01.h>
#include <Concrete01.h>
#include <ElasticPPMaterial.h>
#include <ElasticMaterial.h>
#include <FiberSection3d.h>
#include <ElasticMaterial.h>
#include <ForceBeamColumn3d.h>
#include <LinearCrdTransf3d.h>
#include <ForceBeamColumn2d.h>
#include <LinearCrdTransf2d.h>
#include <LobattoBeamIntegration.h>

#include <AnalysisModel.h>
#include <PlainHandler.h>
#include <TransformationConstraintHandler.h>
#include <PlainNumberer.h>
#include <DOF_Numberer.h>
#include <RCM.h>
#include <DOF_Group.h>
#include <DOF_GrpIter.h>
#include <FE_Element.h>
#include <FE_EleIter.h>
#include <Graph.h>
// The two general sparse SOEs keep their arrays private and offer no accessor;
// the harness has to read A/colA/rowStartA, so it opens them up for this
// translation unit only (the reference sources are not touched).
#define private public
#define protected public
#include <SparseGenRowLinSOE.h>
#include <SparseGenColLinSOE.h>
#undef private
#undef protected
#include <SparseGenRowLinSolver.h>
#include <SparseGenColLinSolver.h>
#include <Beam2dUniformLoad.h>
#include <Beam3dUniformLoad.h>
#include <Beam2dPointLoad.h>
#include <Beam2dPartialUniformLoad.h>
#include <Beam3dPartialUniformLoad.h>
#include <Beam3dPointLoad.h>
#include <BandGenLinSOE.h>
#include <BandGenLinSolver.h>
#include <ProfileSPDLinSOE.h>
#include <ProfileSPDLinSolver.h>
#include <LoadControl.h>
#include <Newmark.h>
#include <TransientIntegrator.h>
#include <DisplacementControl.h>
#include <NewtonRaphson.h>
#include <CTestNormDispIncr.h>
#include <CTestNormUnbalance.h>
#include <CTestEnergyIncr.h>

// ---------------------------------------------------------------------------
// A dense LU solver for SparseGenRowLinSOE.  The linear solve is outside the
// hot path (BASELINE.json: "left to the reference's SOE solver and timed
// separately"); the reference's own SparseGenRow solvers need SuperLU/PETSc,
// so the harness supplies the simplest exact one.
// ---------------------------------------------------------------------------
static int dense_solve(int n, const int* ptr, const int* idx, const double* A, bool csc,
                       const double* Bv, double* X) {
  std::vector<double> M((size_t)n * n, 0.0), b(n);
  for (int r = 0; r < n; r++) {
    for (int k = ptr[r]; k < ptr[r + 1]; k++) {
      if (csc) M[(size_t)idx[k] * n + r] += A[k];
      else     M[(size_t)r * n + idx[k]] += A[k];
    }
    b[r] = Bv[r];
  }
  for (int c = 0; c < n; c++) {
    int p = c; double best = std::fabs(M[(size_t)c * n + c]);
    for (int r = c + 1; r < n; r++)
      if (std::fabs(M[(size_t)r * n + c]) > best) { best = std::fabs(M[(size_t)r * n + c]); p = r; }
    if (best == 0.0) return -1;
    if (p != c) {
      for (int k = 0; k < n; k++) std::swap(M[(size_t)p * n + k], M[(size_t)c * n + k]);
      std::swap(b[p], b[c]);
    }
    for (int r = c + 1; r < n; r++) {
      double f = M[(size_t)r * n + c] / M[(size_t)c * n + c];
      if (f == 0.0) continue;
      for (int k = c; k < n; k++) M[(size_t)r * n + k] -= f * M[(size_t)c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int k = r + 1; k < n; k++) s -= M[(size_t)r * n + k] * X[k];
    X[r] = s / M[(size_t)r * n + r];
  }
  return 0;
}

class HarnessRowSolver : public SparseGenRowLinSolver {
public:
  HarnessRowSolver() : SparseGenRowLinSolver(0) {}
  int setSize() override { return 0; }
  int solve() override {
    return dense_solve(theSOE->size, theSOE->rowStartA, theSOE->colA, theSOE->A, false,
                       &theSOE->B[0], &theSOE->X[0]);
  }
  int sendSelf(int, Channel&) override { return 0; }
  int recvSelf(int, Channel&, FEM_ObjectBroker&) override { return 0; }
};

class HarnessColSolver : public SparseGenColLinSolver {
public:
  HarnessColSolver() : SparseGenColLinSolver(0) {}
  int setSize() override { return 0; }
  int solve() override {
    return dense_solve(theSOE->size, theSOE->colStartA, theSOE->rowA, theSOE->A, true,
                       &theSOE->B[0], &theSOE->X[0]);
  }
  int sendSelf(int, Channel&) override { return 0; }
  int recvSelf(int, Channel&, FEM_ObjectBroker&) override { return 0; }
};

// `system BandGeneral`: the reference's own BandGenLinSOE (its addA and storage are on the path); its LAPACK solver
// (dgbsv) has no library here, so the solve -- outside the path -- expands the band and takes the dense route above.
class BandSOE : public BandGenLinSOE {
 public:
  explicit BandSOE(BandGenLinSolver& s) : BandGenLinSOE(s) {}
  int sub() const { return numSubD; } int super() const { return numSuperD; } int n() const { return size; }
  double* a() { return A; } double* b() { return B; } double* x() { return X; } Vector& bvec() { return *vectB; }
  long long asize() const { return (long long)size * (2 * numSubD + numSuperD + 1); }
};
class HarnessBandSolver : public BandGenLinSolver {
 public:
  HarnessBandSolver() : BandGenLinSolver(0) {}
  int setSize() override { return 0; }
  int solve() override {
    BandSOE* S = static_cast<BandSOE*>(theSOE);
    const int n = S->n(), kl = S->sub(), ku = S->super(), ldA = 2 * kl + ku + 1;
    std::vector<int> ptr(n + 1), idx; std::vector<double> val;
    for (int col = 0; col < n; col++) {            // BandGenLinSOE.cpp:208: (row, col) at A[col ldA + kl + ku - (col - row)]
      ptr[col] = (int)idx.size();
      for (int row = std::max(0, col - ku); row <= std::min(n - 1, col + kl); row++) { idx.push_back(row); val.push_back(S->a()[(size_t)col * ldA + kl + ku - (col - row)]); }
    }
    ptr[n] = (int)idx.size();
    return dense_solve(n, ptr.data(), idx.data(), val.data(), true, S->b(), S->x());
  }
  int sendSelf(int, Channel&) override { return 0; }
  int recvSelf(int, Channel&, FEM_ObjectBroker&) override { return 0; }
};

// SparseGenRowLinSOE leaves LinearSOE::setX pure (its definitions sit under
// "#if 0" in SparseGenRowLinSOE.h:66); give it the obvious one so it can be built.
class RowSOE : public SparseGenRowLinSOE {
public:
  RowSOE(SparseGenRowLinSolver& s) : SparseGenRowLinSOE(s) {}
  void setX(int loc, double value) override { if (loc >= 0 && loc < size) X[loc] = value; }
  void setX(const Vector& x) override { X = x; }
};

struct RefModel {
  int ndm, ndf;
  Domain* domain = nullptr;
  std::map<int, NDMaterial*> ndmats;
  std::map<int, UniaxialMaterial*> unimats;
  std::map<int, SectionForceDeformation*> sections2d;   // FiberSection2d or SectionAggregator
  std::map<int, FiberSection3d*> sections3d;
  AnalysisModel* amodel = nullptr;
  ConstraintHandler* handler = nullptr;
  int handler_kind = 0;          // `constraints Plain` (0) | `constraints Transformation` (1)
  DOF_Numberer* numberer = nullptr;
  LinearSOE* soe = nullptr;
  SparseGenRowLinSOE* rsoe = nullptr;   // soeKind 1
  SparseGenColLinSOE* csoe = nullptr;   // soeKind 0
  BandSOE* bsoe = nullptr;              // soeKind 2 (`system BandGeneral`): no sparse pattern, ptr() / idx() do not apply
  int cur_pattern = 1;                  // the load pattern ref_add_load fills
  double beam_rho = 0.0;                // `-mass` of the forceBeamColumn elements being added
  double beam_off[6] = {0, 0, 0, 0, 0, 0};   // joint offsets of the elements being added (I: 0..2, J: 3..5)
  int n() const { return bsoe ? bsoe->n() : (rsoe ? rsoe->size : csoe->size); }
  const int* ptr() const { return rsoe ? rsoe->rowStartA : csoe->colStartA; }
  const int* idx() const { return rsoe ? rsoe->colA : csoe->rowA; }
  const double* A() const { return bsoe ? bsoe->a() : (rsoe ? rsoe->A : csoe->A); }
  IncrementalIntegrator* integ = nullptr;
  StaticIntegrator* sinteg = nullptr;
  TransientIntegrator* tinteg = nullptr;
  ConvergenceTest* test = nullptr;
  NewtonRaphson* algo = nullptr;
  int nloads = 0;
};

extern "C" {

void* ref_model_new(int ndm, int ndf) {
  RefModel* m = new RefModel;
  m->ndm = ndm; m->ndf = ndf;
  m->domain = new Domain();
  return m;
}

int ref_add_node(void* h, int tag, const double* x) {
  RefModel* m = (RefModel*)h;
  Node* n = (m->ndm == 2) ? new Node(tag, m->ndf, x[0], x[1])
                          : new Node(tag, m->ndf, x[0], x[1], x[2]);
  return m->domain->addNode(n) ? 0 : -1;
}

// a node created under another `model -ndf` (e.g. a FourNodeQuad's 2-dof node in a model whose frame nodes have 3)
int ref_add_node_ndf(void* h, int tag, const double* x, int ndf) {
  RefModel* m = (RefModel*)h;
  Node* n = (m->ndm == 2) ? new Node(tag, ndf, x[0], x[1]) : new Node(tag, ndf, x[0], x[1], x[2]);
  return m->domain->addNode(n) ? 0 : -1;
}

int ref_fix(void* h, int nodeTag, int dof) {
  RefModel* m = (RefModel*)h;
  SP_Constraint* sp = new SP_Constraint(nodeTag, dof, 0.0, true);
  return m->domain->addSP_Constraint(sp) ? 0 : -1;
}

// `equalDOF rNode cNode dofs...` (runtime/commands/domain/constraint.cpp: MP_Constraint with an identity Ccr)
int ref_equal_dof(void* h, int rTag, int cTag, int n, const int* dofs) {
  RefModel* m = (RefModel*)h;
  Matrix Ccr(n, n);
  ID rcDOF(n);
  for (int i = 0; i < n; i++) { Ccr(i, i) = 1.0; rcDOF(i) = dofs[i]; }
  MP_Constraint* mp = new MP_Constraint(rTag, cTag, Ccr, rcDOF, rcDOF);
  return m->domain->addMP_Constraint(mp) ? 0 : -1;
}

// kind 0: ElasticIsotropic (E, nu, rho)        -- runtime/commands/modeling/nDMaterial.cpp
// kind 1: J2Plasticity (K,G,sig0,sigInf,delta,H,eta) -- commands/modeling/material/plastic.cpp:927
int ref_add_nd_material(void* h, int tag, int kind, const double* p) {
  RefModel* m = (RefModel*)h;
  NDMaterial* mat = nullptr;
  if (kind == 0) mat = new ElasticIsotropicMaterial(tag, p[0], p[1], p[2]);
  else if (kind == 1) mat = new J2Plasticity(tag, 0, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);   // p[7] = rho
  if (!mat) return -1;
  m->ndmats[tag] = mat;
  return 0;
}

int ref_add_brick(void* h, int tag, const int* nd, int matTag, const double* b) {
  RefModel* m = (RefModel*)h;
  Element* e = new Brick(tag, nd[0], nd[1], nd[2], nd[3], nd[4], nd[5], nd[6], nd[7],
                         *m->ndmats.at(matTag), b[0], b[1], b[2]);
  return m->domain->addElement(e) ? 0 : -1;
}

// type: 0 PlaneStrain, 1 PlaneStress
int ref_add_quad(void* h, int tag, const int* nd, int matTag, double thick, int type,
                 double pressure, double rho, const double* b) {
  RefModel* m = (RefModel*)h;
  NDMaterial* copy = m->ndmats.at(matTag)->getCopy(type == 0 ? "PlaneStrain" : "PlaneStress");
  if (!copy) return -2;
  std::array<int, 4> nodes{nd[0], nd[1], nd[2], nd[3]};
  Element* e = new FourNodeQuad(tag, nodes, *copy, thick, pressure, rho, b[0], b[1]);
  delete copy;
  return m->domain->addElement(e) ? 0 : -1;
}

// kind 0: Steel02 (Fy,E0,b,R0,cR1,cR2,a1,a2,a3,a4,sigInit); kind 1: Concrete02 (fc,epsc0,fcu,epscu,rat,ft,Ets);
// kind 2: Steel01 (fy,E0,b,a1,a2,a3,a4); kind 3: Elastic (E,eta,Eneg)
static UniaxialMaterial* make_uniaxial(int tag, int kind, const double* p) {
  if (kind == 2) return new Steel01(tag, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
  if (kind == 3) return new ElasticMaterial(tag, p[0], p[1], p[2]);
  if (kind == 4) return new Concrete01(tag, p[0], p[1], p[2], p[3]);          // fpc, epsc0, fpcu, epscu
  if (kind == 5) return new ElasticPPMaterial(tag, p[0], p[1], p[2], p[3]);   // E, epsyP, epsyN, eps0
  if (kind == 0) return new Steel02(tag, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]);
  if (kind == 1) return new Concrete02(tag, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
  return nullptr;
}
int ref_add_uniaxial(void* h, int tag, int kind, const double* p) {
  RefModel* m = (RefModel*)h;
  UniaxialMaterial* u = make_uniaxial(tag, kind, p);
  if (!u) return -1;
  m->unimats[tag] = u;
  return 0;
}
// section Fiber (FiberSection2d, centroid computed as the section command does by default)
int ref_add_fiber_section(void* h, int tag, int nf, const double* y, const double* A, const int* matTags) {
  RefModel* m = (RefModel*)h;
  FiberSection2d* s = new FiberSection2d(tag, nf, true);
  for (int i = 0; i < nf; i++)
    if (s->addFiber(*m->unimats.at(matTags[i]), A[i], y[i]) < 0) return -1;
  m->sections2d[tag] = s;
  return 0;
}
// section Aggregator tag mat1 code1 mat2 code2 ... (no base section): SectionAggregator.cpp:119
int ref_add_section_aggregator(void* h, int tag, int n, const int* matTags, const int* codes) {
  RefModel* m = (RefModel*)h;
  std::vector<UniaxialMaterial*> adds(n);
  ID code(n);
  for (int i = 0; i < n; i++) { adds[i] = m->unimats.at(matTags[i]); code(i) = codes[i]; }
  m->sections2d[tag] = new SectionAggregator(tag, n, adds.data(), code);
  return 0;
}
// element forceBeamColumn (2D): Lobatto integration, Linear transformation, nIP copies of one section
// -integration: 0 Lobatto, 1 Legendre, 2 Radau, 3 NewtonCotes, 4 Trapezoidal
static BeamIntegration* make_beam_integration(int kind) {
  switch (kind) {
    case 1: return new LegendreBeamIntegration();
    case 2: return new RadauBeamIntegration();
    case 3: return new NewtonCotesBeamIntegration();
    case 4: return new TrapezoidalBeamIntegration();
    default: return new LobattoBeamIntegration();
  }
}
// what the rule gives for nip sections on an element of length L: locations and weights as fractions of L
int ref_beam_rule(int kind, int nip, double L, double* xi, double* wt) {
  BeamIntegration* bi = make_beam_integration(kind);
  bi->getSectionLocations(nip, L, xi);
  bi->getSectionWeights(nip, L, wt);
  delete bi;
  return 0;
}
int ref_add_force_beam2d_t(void* h, int tag, const int* nd, int secTag, int nip, int maxIters, double tol, int transfKind) {
  RefModel* m = (RefModel*)h;
  std::vector<SectionForceDeformation*> secs(nip, m->sections2d.at(secTag));
  BeamIntegration* bip = make_beam_integration(transfKind / 16); transfKind %= 16;      // (integration kind in the upper bits)
  BeamIntegration& bi = *bip;
  Vector oI(2), oJ(2);                        // -jntOffset dXi dYi dXj dYj
  oI(0) = m->beam_off[0]; oI(1) = m->beam_off[1]; oJ(0) = m->beam_off[3]; oJ(1) = m->beam_off[4];
  const bool off = oI(0) != 0.0 || oI(1) != 0.0 || oJ(0) != 0.0 || oJ(1) != 0.0;
  LinearCrdTransf2d lin0(tag), lin1(tag, oI, oJ);
  PDeltaCrdTransf2d pd0(tag), pd1(tag, oI, oJ);      // geomTransf PDelta
  LinearCrdTransf2d& lin = off ? lin1 : lin0;
  PDeltaCrdTransf2d& pd = off ? pd1 : pd0;
  CorotCrdTransf2d cor(tag, oI, oJ);                 // geomTransf Corotational
  CrdTransf& transf = transfKind == 2 ? (CrdTransf&)cor : (transfKind == 1 ? (CrdTransf&)pd : (CrdTransf&)lin);
  Element* e = new ForceBeamColumn2d(tag, nd[0], nd[1], nip, secs.data(), bi, transf, m->beam_rho, maxIters, tol);
  return m->domain->addElement(e) ? 0 : -1;
}
// `geomTransf ... -jntOffset`: offsets (global components, I then J, three values each) of the elements added from now on
int ref_set_beam_offsets(void* h, const double* off6) { for (int i = 0; i < 6; i++) ((RefModel*)h)->beam_off[i] = off6[i]; return 0; }
// `-mass rho` of the forceBeamColumn elements added from now on
int ref_set_beam_rho(void* h, double rho) { ((RefModel*)h)->beam_rho = rho; return 0; }
int ref_add_force_beam2d(void* h, int tag, const int* nd, int secTag, int nip, int maxIters, double tol) {
  return ref_add_force_beam2d_t(h, tag, nd, secTag, nip, maxIters, tol, 0);
}

// section Fiber tag -GJ gj { fiber y z A mat ... } in a 3D model (runtime/commands/modeling/section.cpp:497):
// FiberSection3d with an elastic torsion material, centroid computed
int ref_add_fiber_section3d(void* h, int tag, int nf, const double* y, const double* z, const double* A, const int* matTags, double GJ) {
  RefModel* m = (RefModel*)h;
  ElasticMaterial torsion(0, GJ);
  FiberSection3d* s = new FiberSection3d(tag, nf, torsion, true);
  for (int i = 0; i < nf; i++)
    if (s->addFiber(*m->unimats.at(matTags[i]), A[i], y[i], z[i]) < 0) return -1;
  m->sections3d[tag] = s;
  return 0;
}
// element forceBeamColumn (3D, frames.cpp:333 -> ForceBeamColumn3d): Lobatto integration,
// geomTransf Linear with vecxz, nIP copies of one section
int ref_add_force_beam3d_t(void* h, int tag, const int* nd, int secTag, int nip, int maxIters, double tol, const double* vecxz, int transfKind) {
  RefModel* m = (RefModel*)h;
  std::vector<SectionForceDeformation*> secs(nip, m->sections3d.at(secTag));
  BeamIntegration* bip = make_beam_integration(transfKind / 16); transfKind %= 16;
  BeamIntegration& bi = *bip;
  Vector v(3); v(0) = vecxz[0]; v(1) = vecxz[1]; v(2) = vecxz[2];
  Vector oI(3), oJ(3);                        // -jntOffset dXi dYi dZi dXj dYj dZj
  for (int q = 0; q < 3; q++) { oI(q) = m->beam_off[q]; oJ(q) = m->beam_off[3 + q]; }
  const bool off = oI.Norm() != 0.0 || oJ.Norm() != 0.0;
  LinearCrdTransf3d lin0(tag, v), lin1(tag, v, oI, oJ);
  PDeltaCrdTransf3d pd0(tag, v), pd1(tag, v, oI, oJ);      // geomTransf PDelta
  LinearCrdTransf3d& lin = off ? lin1 : lin0;
  PDeltaCrdTransf3d& pd = off ? pd1 : pd0;
  CrdTransf& transf = transfKind == 1 ? (CrdTransf&)pd : (CrdTransf&)lin;
  Element* e = new ForceBeamColumn3d(tag, nd[0], nd[1], nip, secs.data(), bi, transf, m->beam_rho, maxIters, tol);
  return m->domain->addElement(e) ? 0 : -1;
}
int ref_add_force_beam3d(void* h, int tag, const int* nd, int secTag, int nip, int maxIters, double tol, const double* vecxz) {
  return ref_add_force_beam3d_t(h, tag, nd, secTag, nip, maxIters, tol, vecxz, 0);
}

int ref_uni_path(int kind, const double* p, int n, const double* strains, const int* commit,
                 double* stress, double* tangent) {
  UniaxialMaterial* mat = make_uniaxial(1, kind, p);
  if (!mat) return -1;
  for (int s = 0; s < n; s++) {
    if (mat->setTrialStrain(strains[s]) < 0) return -2;
    stress[s] = mat->getStress(); tangent[s] = mat->getTangent();
    if (commit[s]) mat->commitState();
  }
  delete mat;
  return 0;
}

// `eleLoad -ele tag -type -beamUniform wy [wz] wa` in pattern 1 (Linear series)
int ref_add_beam_uniform_load(void* h, int eleTag, double wy, double wz, double wa) {
  RefModel* m = (RefModel*)h;
  if (m->domain->getLoadPattern(m->cur_pattern) == nullptr) {
    LoadPattern* lp = new LoadPattern(m->cur_pattern);
    lp->setTimeSeries(new LinearSeries());
    m->domain->addLoadPattern(lp);
  }
  ElementalLoad* el = (m->ndm == 2) ? (ElementalLoad*)new Beam2dUniformLoad(10000 + m->nloads++, wy, wa, eleTag)
                                    : (ElementalLoad*)new Beam3dUniformLoad(10000 + m->nloads++, wy, wz, wa, eleTag);
  return m->domain->addElementalLoad(el, 1) ? 0 : -1;
}

// `eleLoad -ele tag -type -beamPoint Py [Pz] xL [N]` in pattern 1
int ref_add_beam_point_load(void* h, int eleTag, double Py, double Pz, double N, double aOverL) {
  RefModel* m = (RefModel*)h;
  if (m->domain->getLoadPattern(m->cur_pattern) == nullptr) {
    LoadPattern* lp = new LoadPattern(m->cur_pattern);
    lp->setTimeSeries(new LinearSeries());
    m->domain->addLoadPattern(lp);
  }
  ElementalLoad* el = (m->ndm == 2) ? (ElementalLoad*)new Beam2dPointLoad(20000 + m->nloads++, Py, aOverL, eleTag, N)
                                    : (ElementalLoad*)new Beam3dPointLoad(20000 + m->nloads++, Py, Pz, aOverL, eleTag, N);
  return m->domain->addElementalLoad(el, 1) ? 0 : -1;
}

// `eleLoad -ele tag -type -beamUniform wya wyb ... aOverL bOverL` (2D, trapezoidal over part of the element) in pattern 1
int ref_add_beam_partial_load(void* h, int eleTag, const double* q) {
  RefModel* m = (RefModel*)h;
  if (m->domain->getLoadPattern(m->cur_pattern) == nullptr) {
    LoadPattern* lp = new LoadPattern(m->cur_pattern);
    lp->setTimeSeries(new LinearSeries());
    m->domain->addLoadPattern(lp);
  }
  // q = wy_a, wy_b, wAxial_a, wAxial_b, aOverL, bOverL, wz_a, wz_b
  ElementalLoad* el = (m->ndm == 2) ? (ElementalLoad*)new Beam2dPartialUniformLoad(30000 + m->nloads++, q[0], q[1], q[2], q[3], q[4], q[5], eleTag)
                                    : (ElementalLoad*)new Beam3dPartialUniformLoad(30000 + m->nloads++, q[0], q[6], q[2], q[4], q[5], q[1], q[7], q[3], eleTag);
  return m->domain->addElementalLoad(el, 1) ? 0 : -1;
}

int ref_add_load(void* h, int nodeTag, const double* vals) {
  RefModel* m = (RefModel*)h;
  if (m->domain->getLoadPattern(m->cur_pattern) == nullptr) {
    LoadPattern* lp = new LoadPattern(m->cur_pattern);
    lp->setTimeSeries(new LinearSeries());
    m->domain->addLoadPattern(lp);
  }
  Node* node = m->domain->getNode(nodeTag);
  const int nd = node ? node->getNumberDOF() : m->ndf;       // (a node may carry fewer dofs than the model's ndf)
  Vector v(nd);
  for (int i = 0; i < nd; i++) v(i) = vals[i];
  NodalLoad* nl = new NodalLoad(m->nloads++, nodeTag, v);
  return m->domain->addNodalLoad(nl, m->cur_pattern) ? 0 : -1;
}
// `loadConst -time t` (runtime/commands/domain/...: Domain::setLoadConstant, setCurrentTime, setCommittedTime); the next
// ref_add_load opens a new `pattern Plain n Linear`
int ref_load_const(void* h, double time) {
  RefModel* m = (RefModel*)h;
  m->domain->setLoadConstant();
  m->domain->setCurrentTime(time);
  m->domain->setCommittedTime(time);
  m->cur_pattern++;
  return 0;
}

// restates BasicAnalysisBuilder::domainChanged (BasicAnalysisBuilder.cpp:225-300)
// numberer: 0 Plain, 1 RCM.  integrator: LoadControl(dlambda).
// soeKind: 0 SparseGenColLinSOE (CSC, the live "SparseGeneral" system), 1 SparseGenRowLinSOE (CSR)
static int ref_setup_common(RefModel* m, int numberer, int soeKind, int testKind, double tol, int maxIter);

int ref_setup(void* h, int numberer, int soeKind, double dlambda, int testKind, double tol, int maxIter) {
  RefModel* m = (RefModel*)h;
  m->sinteg = new LoadControl(dlambda, 1, dlambda, dlambda);
  m->integ = m->sinteg;
  return ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
}

// integrator DisplacementControl node dof incr (analysis/integrator/Static/DisplacementControl.cpp);
// dof is 0-based here as inside the class (the Tcl command subtracts 1)
int ref_setup_dispcontrol(void* h, int numberer, int soeKind, int node, int dof, double incr, int testKind, double tol, int maxIter) {
  RefModel* m = (RefModel*)h;
  m->sinteg = new DisplacementControl(node, dof, incr, m->domain, 1, incr, incr);
  m->integ = m->sinteg;
  return ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
}
double ref_get_lambda(void* h) { return ((RefModel*)h)->amodel->getCurrentDomainTime(); }

// `mass` command: Node::setMass with a diagonal matrix (domain/node/Node.h:127)
int ref_set_mass(void* h, int nodeTag, const double* mvals) {
  RefModel* m = (RefModel*)h;
  Node* n = m->domain->getNode(nodeTag);
  if (!n) return -1;
  Matrix M(m->ndf, m->ndf);
  for (int i = 0; i < m->ndf; i++) M(i, i) = mvals[i];
  return n->setMass(M);
}
// rayleigh alphaM 0 0 0 restricted to the nodes (Node::setRayleighDampingFactor)
int ref_set_alphaM(void* h, double alphaM) {
  RefModel* m = (RefModel*)h;
  NodeIter& it = m->domain->getNodes(); Node* n;
  while ((n = it()) != nullptr) n->setRayleighDampingFactor(alphaM);
  return 0;
}
// `rayleigh alphaM betaK betaKinit betaKcomm` (Domain::setRayleighDampingFactors, Domain.cpp:1858)
int ref_set_rayleigh(void* h, double alphaM, double betaK, double betaK0, double betaKc) {
  return ((RefModel*)h)->domain->setRayleighDampingFactors(alphaM, betaK, betaK0, betaKc);
}
// integrator Newmark gamma beta (displacement form); analysis Transient
int ref_setup_transient(void* h, int numberer, int soeKind, double gamma, double beta, int testKind, double tol, int maxIter) {
  RefModel* m = (RefModel*)h;
  m->tinteg = new Newmark(gamma, beta);
  m->integ = m->tinteg;
  return ref_setup_common(m, numberer, soeKind, testKind, tol, maxIter);
}
int ref_transient_new_step(void* h, double dt) { return ((RefModel*)h)->tinteg->newStep(dt); }
int ref_transient_update(void* h, const double* dU) {
  RefModel* m = (RefModel*)h;
  Vector v(m->n());
  for (int i = 0; i < m->n(); i++) v(i) = dU[i];
  return m->tinteg->update(v);
}
int ref_get_vel_accel(void* h, int n, const int* tags, double* vel, double* acc) {
  RefModel* m = (RefModel*)h;
  for (int i = 0; i < n; i++) {
    Node* nd = m->domain->getNode(tags[i]);
    const Vector& v = nd->getTrialVel(); const Vector& a = nd->getTrialAccel();
    for (int j = 0; j < m->ndf; j++) { vel[(size_t)i * m->ndf + j] = v(j); acc[(size_t)i * m->ndf + j] = a(j); }
  }
  return 0;
}
// BasicAnalysisBuilder::analyzeStep (BasicAnalysisBuilder.cpp:464): newStep(dT), solveCurrentStep, commit
int ref_analyze_transient(void* h, int nsteps, double dt, int* iters, double* norms, int maxIter) {
  RefModel* m = (RefModel*)h;
  for (int s = 0; s < nsteps; s++) {
    if (m->amodel->analysisStep(dt) < 0) return -2;
    if (m->tinteg->newStep(dt) < 0) return -2;
    int r = m->algo->solveCurrentStep();
    iters[s] = m->test->getNumTests();
    if (norms) {
      const Vector& nv = m->test->getNorms();
      for (int k = 0; k < maxIter && k < nv.Size(); k++) norms[(size_t)s * maxIter + k] = nv(k);
    }
    if (r < 0) { m->domain->revertToLastCommit(); m->tinteg->revertToLastStep(); return -3; }
    if (m->tinteg->commit() < 0) return -4;
  }
  return 0;
}

static int ref_setup_common(RefModel* m, int numberer, int soeKind, int testKind, double tol, int maxIter) {
  m->amodel = new AnalysisModel();
  m->handler = m->handler_kind == 1 ? (ConstraintHandler*)new TransformationConstraintHandler() : (ConstraintHandler*)new PlainHandler();
  if (numberer == 0) m->numberer = new PlainNumberer();
  else { RCM* rcm = new RCM(false); m->numberer = new DOF_Numberer(*rcm); }
  m->rsoe = nullptr; m->csoe = nullptr; m->bsoe = nullptr;     // (a second `analysis` on the same Domain starts over)
  if (soeKind == 2) { m->bsoe = new BandSOE(*new HarnessBandSolver()); m->soe = m->bsoe; }
  else if (soeKind == 1) { m->rsoe = new RowSOE(*new HarnessRowSolver()); m->soe = m->rsoe; }
  else { m->csoe = new SparseGenColLinSOE(*new HarnessColSolver()); m->soe = m->csoe; }
  if (testKind == 0) m->test = new CTestNormDispIncr(tol, maxIter, 0);
  else if (testKind == 1) m->test = new CTestNormUnbalance(tol, maxIter, 0);
  else m->test = new CTestEnergyIncr(tol, maxIter, 0);
  m->algo = new NewtonRaphson(*m->test);

  m->amodel->setLinks(*m->domain, *m->handler);
  m->handler->setLinks(*m->domain, *m->amodel, *m->integ);
  m->numberer->setLinks(*m->amodel);
  m->integ->setLinks(*m->amodel, *m->soe, m->test);
  m->algo->setLinks(*m->amodel, *m->integ, *m->soe, m->test);
  m->soe->setLinks(*m->amodel);

  m->amodel->clearAll();
  m->handler->clearAll();
  if (m->handler->handle() < 0) return -1;
  if (m->numberer->numberDOF() < 0) return -2;
  if (m->handler->doneNumberingDOF() < 0) return -2;
  Graph& g = m->amodel->getDOFGraph();
  if (m->soe->setSize(g) < 0) return -3;
  m->amodel->clearDOFGraph();
  if (m->integ->domainChanged() < 0) return -4;
  return m->amodel->getNumEqn();
}

// `constraints Plain | Transformation` of the analyses set up from now on
int ref_set_handler(void* h, int kind) { ((RefModel*)h)->handler_kind = kind; return 0; }

int ref_num_eqn(void* h) { return ((RefModel*)h)->soe->getNumEqn(); }
int ref_nnz(void* h) { RefModel* m = (RefModel*)h; return m->bsoe ? (int)m->bsoe->asize() : m->ptr()[m->n()]; }

// equation ids of a node's DOF_Group (DOF_Group::getID)
int ref_node_ids(void* h, int nodeTag, int* ids) {
  RefModel* m = (RefModel*)h;
  Node* n = m->domain->getNode(nodeTag);
  if (!n || !n->getDOF_GroupPtr()) return -1;
  const ID& id = n->getDOF_GroupPtr()->getID();
  for (int i = 0; i < id.Size(); i++) ids[i] = id(i);
  return id.Size();
}

// FE_Element::getID in AnalysisModel iteration order; returns count of FE elements
int ref_fe_ids(void* h, int* eleTags, int* ids, int stride) {
  RefModel* m = (RefModel*)h;
  FE_EleIter& it = m->amodel->getFEs();
  FE_Element* fe; int c = 0;
  while ((fe = it()) != nullptr) {
    const ID& id = fe->getID();
    if (eleTags) eleTags[c] = fe->getTag();
    for (int i = 0; i < id.Size() && i < stride; i++) ids[(size_t)c * stride + i] = id(i);
    c++;
  }
  return c;
}

void ref_get_csr(void* h, int* rowStart, int* colA) {
  RefModel* m = (RefModel*)h;
  int n = m->n();
  memcpy(rowStart, m->ptr(), sizeof(int) * (n + 1));
  memcpy(colA, m->idx(), sizeof(int) * m->ptr()[n]);
}

// Node::setTrialDisp for every node (u is [nNodes][ndf] in the given tag order),
// then Domain::update -> Element::update (state determination)
int ref_set_trial_disp(void* h, int n, const int* tags, const double* u) {
  RefModel* m = (RefModel*)h;
  for (int i = 0; i < n; i++) {
    Node* nd = m->domain->getNode(tags[i]);
    Vector v(nd->getNumberDOF());
    for (int j = 0; j < v.Size(); j++) v(j) = u[(size_t)i * m->ndf + j];
    nd->setTrialDisp(v);
  }
  return m->domain->update();
}

int ref_get_trial_disp(void* h, int n, const int* tags, double* u) {
  RefModel* m = (RefModel*)h;
  for (int i = 0; i < n; i++) {
    const Vector& d = m->domain->getNode(tags[i])->getTrialDisp();
    for (int j = 0; j < m->ndf; j++) u[(size_t)i * m->ndf + j] = j < d.Size() ? d(j) : 0.0;
  }
  return 0;
}

void ref_apply_load(void* h, double lambda) { ((RefModel*)h)->amodel->applyLoadDomain(lambda); }

// IncrementalIntegrator::formTangent (IncrementalIntegrator.cpp:74) -> A values (CSR order)
int ref_form_tangent(void* h, double* A) {
  RefModel* m = (RefModel*)h;
  int r = m->integ->formTangent(CURRENT_TANGENT);
  if (A) memcpy(A, m->A(), sizeof(double) * ref_nnz(h));
  return r;
}

// `system BandGeneral` / `system ProfileSPD` of the SAME analysis: the reference's own BandGenLinSOE (kind 2) or
// ProfileSPDLinSOE (kind 3) is sized from the AnalysisModel's DOF graph and filled the way
// IncrementalIntegrator::formTangent does for a static integrator (zeroA; addA(FE_Element::getTangent, getID) in
// FE_Element order).  layout: kind 2 -> {numSubD, numSuperD}, kind 3 -> iDiagLoc[numEqn].  Returns the length of A.
namespace {
class NoBandSolver : public BandGenLinSolver {
 public:
  NoBandSolver() : BandGenLinSolver(0) {}
  int solve() override { return 0; }
  int setSize() override { return 0; }
  int sendSelf(int, Channel&) override { return 0; }
  int recvSelf(int, Channel&, FEM_ObjectBroker&) override { return 0; }
};
class NoProfileSolver : public ProfileSPDLinSolver {
 public:
  NoProfileSolver() : ProfileSPDLinSolver(0) {}
  int solve() override { return 0; }
  int setSize() override { return 0; }
  int sendSelf(int, Channel&) override { return 0; }
  int recvSelf(int, Channel&, FEM_ObjectBroker&) override { return 0; }
};
struct BandPeek : public BandGenLinSOE {
  explicit BandPeek(BandGenLinSolver& s) : BandGenLinSOE(s) {}
  int sub() const { return numSubD; } int super() const { return numSuperD; } int n() const { return size; } const double* a() const { return A; }
};
struct ProfilePeek : public ProfileSPDLinSOE {
  explicit ProfilePeek(ProfileSPDLinSolver& s) : ProfileSPDLinSOE(s) {}
  int n() const { return size; } int psize() const { return profileSize; } const int* diag() const { return iDiagLoc; } const double* a() const { return A; }
};
}  // namespace
long long ref_store_tangent(void* h, int kind, int* layout, double* A) {
  RefModel* m = (RefModel*)h;
  Graph& g = m->amodel->getDOFGraph();
  long long n = -1;
  auto fill = [&](LinearSOE& soe) {
    soe.zeroA();
    FE_EleIter& eles = m->amodel->getFEs();
    FE_Element* fe;
    while ((fe = eles()) != nullptr) soe.addA(fe->getTangent(m->integ), fe->getID());
  };
  if (kind == 2) {
    BandPeek soe(*new NoBandSolver());
    if (soe.setSize(g) < 0) return -1;
    fill(soe);
    layout[0] = soe.sub(); layout[1] = soe.super();
    n = (long long)soe.n() * (2 * soe.sub() + soe.super() + 1);
    if (A) memcpy(A, soe.a(), sizeof(double) * n);
  } else if (kind == 3) {
    ProfilePeek soe(*new NoProfileSolver());
    if (soe.setSize(g) < 0) return -1;
    fill(soe);
    for (int i = 0; i < soe.n(); i++) layout[i] = soe.diag()[i];
    n = soe.psize();
    if (A) memcpy(A, soe.a(), sizeof(double) * n);
  }
  m->amodel->clearDOFGraph();
  return n;
}

// IncrementalIntegrator::formUnbalance -> B
int ref_form_unbalance(void* h, double* B) {
  RefModel* m = (RefModel*)h;
  int r = m->integ->formUnbalance();
  if (B) { const Vector& b = m->soe->getB(); for (int i = 0; i < b.Size(); i++) B[i] = b(i); }
  return r;
}

int ref_commit(void* h) { return ((RefModel*)h)->domain->commit(); }
int ref_revert(void* h) { return ((RefModel*)h)->domain->revertToLastCommit(); }
int ref_revert_to_start(void* h) { return ((RefModel*)h)->domain->revertToStart(); }

// element level: Element::getTangentStiff / getResistingForce
int ref_ele_tangent(void* h, int tag, double* K) {
  Element* e = ((RefModel*)h)->domain->getElement(tag);
  if (!e) return -1;
  const Matrix& k = e->getTangentStiff();
  for (int i = 0; i < k.noRows(); i++)
    for (int j = 0; j < k.noCols(); j++) K[(size_t)i * k.noCols() + j] = k(i, j);
  return k.noRows();
}
int ref_ele_resid(void* h, int tag, double* R) {
  Element* e = ((RefModel*)h)->domain->getElement(tag);
  if (!e) return -1;
  const Vector& r = e->getResistingForce();
  for (int i = 0; i < r.Size(); i++) R[i] = r(i);
  return r.Size();
}

// restates BasicAnalysisBuilder::analyzeStatic (BasicAnalysisBuilder.cpp:337-420):
// newStep / solveCurrentStep / commit; iters[i] = Newton iterations of step i
// (ConvergenceTest::getNumTests), norms[i*maxIter + k] = test norms.
int ref_analyze_static_lam(void* h, int nsteps, int* iters, double* norms, int maxIter, double* lam);
int ref_analyze_static(void* h, int nsteps, int* iters, double* norms, int maxIter) {
  return ref_analyze_static_lam(h, nsteps, iters, norms, maxIter, nullptr);
}
int ref_analyze_static_lam(void* h, int nsteps, int* iters, double* norms, int maxIter, double* lam) {
  RefModel* m = (RefModel*)h;
  for (int s = 0; s < nsteps; s++) {
    if (m->amodel->analysisStep(0.0) < 0) return -2;
    if (m->sinteg->newStep() < 0) return -2;
    int r = m->algo->solveCurrentStep();
    iters[s] = m->test->getNumTests();
    if (lam) lam[s] = m->amodel->getCurrentDomainTime();
    if (norms) {
      const Vector& nv = m->test->getNorms();
      for (int k = 0; k < maxIter && k < nv.Size(); k++) norms[(size_t)s * maxIter + k] = nv(k);
    }
    if (r < 0) { m->domain->revertToLastCommit(); m->sinteg->revertToLastStep(); return -3; }
    if (m->sinteg->commit() < 0) return -4;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// material level: strain path through NDMaterial (setTrialStrain / getStress /
// getTangent / commitState).  type: "ThreeDimensional" | "PlaneStrain" ...
// strains [n][order]; commit[n] flags; out stress [n][order], tangent [n][order*order]
// ---------------------------------------------------------------------------
int ref_nd_path(int kind, const double* p, const char* type, int n, const double* strains,
                const int* commit, double* stress, double* tangent) {
  NDMaterial* base = nullptr;
  if (kind == 0) base = new ElasticIsotropicMaterial(1, p[0], p[1], p[2]);
  else base = new J2Plasticity(1, 0, p[0], p[1], p[2], p[3], p[4], p[5], p[6], 0.0);
  NDMaterial* mat = base->getCopy(type);
  if (!mat) return -1;
  int order = mat->getOrder();
  Vector e(order);
  for (int s = 0; s < n; s++) {
    for (int i = 0; i < order; i++) e(i) = strains[(size_t)s * order + i];
    if (mat->setTrialStrain(e) < 0) return -2;
    const Vector& sig = mat->getStress();
    const Matrix& D = mat->getTangent();
    for (int i = 0; i < order; i++) {
      stress[(size_t)s * order + i] = sig(i);
      for (int j = 0; j < order; j++) tangent[((size_t)s * order + i) * order + j] = D(i, j);
    }
    if (commit[s]) mat->commitState();
  }
  delete mat; delete base;
  return order;
}

} // extern "C"
