// Force-based beam-column with fibre sections on the device.
//
// Reference counterparts (under /root/reference/SRC):
//   ForceBeamColumn2d::update / commitState / revertToLastCommit / getTangentStiff /
//     getResistingForce                  element/Frame/Other/Force/ForceBeamColumn2d.cpp:559,276,310,400,523
//   LinearCrdTransf2d (no offsets)       coordTransformation/LinearCrdTransf2d.cpp
//   LobattoBeamIntegration               quadrature/Frame/LobattoBeamIntegration.cpp
//   FiberSection2d::setTrialSectionDeformation / revertToLastCommit
//                                        material/section/FiberSection2d.cpp:225,376
//   SectionForceDeformation::getSectionFlexibility -> cmx_inv2   matrix/routines/invGL2.c
//   Steel02::setTrialStrain              material/uniaxial/steel/Steel02.cpp:113
//   Concrete02::setTrialStrain           material/uniaxial/concrete/Concrete02.cpp:167
//
// One thread per element.  All state is SoA over the element index (coalesced across the
// threads of a warp): element state, per-section state, and per-fibre history with a fixed
// record of XB_FIB_NV doubles (committed and trial copies in separate buffers, so that
// commit / revert are plain device copies).
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace xbk {

constexpr int XB_FIB_NV = 11;   // doubles per fibre record
constexpr int XB_FIB_CURV = 16; // fkind flag: the material of a section Aggregator's Mz response (strain = the curvature)
constexpr int XB_MAXSEC = 10;
constexpr int XB_FBC3D_MAX_PASSES = 20000;
#ifndef XB_FBC_SEC_OCC
#define XB_FBC_SEC_OCC 5
#endif   // see fbc3d_update_kernel
// Steel02 record:   0 epsmin 1 epsmax 2 epspl 3 epss0 4 sigs0 5 epsr 6 sigr 7 kon 8 e 9 sig 10 eps
// Concrete02 record: 0 ecmin 1 dept 8 e 9 sig 10 eps
// Steel01 record:    0 minStrain 1 maxStrain 2 shiftP 3 shiftN 4 loading 8 tangent 9 stress 10 strain
// Elastic record:    8 tangent 9 stress 10 strain
// Concrete01 record: 0 minStrain 1 endStrain 2 unloadSlope 8 tangent 9 stress 10 strain

struct BeamView {
  long long n;                 // elements
  int nip, nf, maxIters;
  double tol;
  const int* conn;             // [n][2]
  const double* geo;           // 2D: [3][n]  L, cosTheta, sinTheta;  3D: [10][n]  L, R[3][3] row-major
  int nb, ord;                 // basic dofs (3 | 6) and section order (2 | 4): sizes of the arrays below
  // section template (shared by every section of every element of the group)
  const double* fy;            // [nf] y - yBar
  const double* fz;            // [nf] z - zBar (3D)
  double GJ;                   // elastic torsion (3D)
  const double* fA;            // [nf]
  const int* fkind;            // [nf] 0 Steel02, 1 Concrete02, 2 Steel01, 3 Elastic, 4 Concrete01 (| XB_FIB_CURV)
  int agg;                     // section Aggregator (P, Mz): flexibility 1/k on the diagonal (SectionAggregator.cpp:419)
  const double* fpar;          // [nf][12] material parameters
  const double* fs0;           // [ord*ord] initial section flexibility (column-major)
  // element state, SoA [k][n]
  double* Se;                  // [nb][n]
  double* kv;                  // [nb*nb][n] column-major
  double* Sec;                 // committed
  double* kvc;
  int* iflag;                  // [n] initialFlag
  // section state [i][k][n]
  double* vs;                  // [nip][ord][n]
  double* fs;                  // [nip][ord*ord][n]
  double* Ssr;                 // [nip][ord][n]
  double* vsc;                 // [nip][ord][n] committed
  // fibre records [ (i*nf+f)*NV + v ][n]
  double* fc;                  // committed
  double* ft;                  // trial
  const long long* kdst;       // [n][2] slot of node a in KeN (or send buffer)
  double* KeN;
  double* sendK;
  int cps;
  double* Re;                  // [n][2*ndf]
  // Rayleigh damping (Element::getDamp): initial stiffness kv0 = inverse of the initial flexibility
  // (getInitialStiff), Kc = kv at the last commit (Element::commitState); both basic, column-major [nb*nb][n]
  double* kv0;
  double* kvK;
  double* nK;                  // [n] the axial force kvK was taken at (the geometric part of Kc under geomTransf PDelta)
  // `eleLoad -beamUniform` of the Linear pattern: wy, wz, wa per element [3][n] (null: none), the pattern's load factor,
  // and whether Domain::applyLoad has run (numEleLoads > 0: the element iterates at every update)
  const double* wl;            // [7][n]: wy, wz, wa, then `eleLoad -beamPoint` Py, Pz, N, aOverL (has_point)
  double lam;
  int loads_on;
  int has_point;
  const int* ulist;            // null, or [nlist]: update only these elements (`constraints Transformation`, see GroupView::ulist)
  long long nlist;
  int has_partial;             // rows 7..14 of wl hold Beam2d/3dPartialUniformLoad's wya, wyb, waa, wab, aOverL, bOverL, wza, wzb
  // geomTransf PDelta (PDeltaCrdTransf2d.cpp / PDeltaCrdTransf3d.cpp): geometric stiffness N/L and leaning-column shear.
  // 2D: the relative transverse displacement is taken from the trial displacements U whenever the element forms its
  // forces (ForceBeamColumn2d.cpp:402,526 refresh the transformation); 3D: ul17, ul28 as of the element's last update
  // (ForceBeamColumn3d never refreshes it), kept in ul [2][n]
  int pdelta;
  // geomTransf Corotational (2D, CorotCrdTransf2d.cpp, no joint offsets): ul = [6][n], the basic deformations ub of the last
  // crdTransf->update() (rows 0..2: the next update's ubpr) and of the last commit (rows 3..5: ubcommit)
  int corot;
  const double* U;
  double* ul;
  // beam integration other than Lobatto (xb_set_beam_integration): [2 nip][n] locations then weights, fractions of L; null: Lobatto
  const double* rule;
  // rigid joint offsets (geomTransf -jntOffset; nodeIOffset / nodeJOffset of Linear / PDeltaCrdTransf2d.cpp, 3d.cpp):
  // 2D [4][n] dXi dYi dXj dYj, 3D [6][n] dXi dYi dZi dXj dYj dZj, null: none.  The element ends follow their nodes rigidly,
  // u_end = u + theta x offset; tangent and forces are pulled back to the nodes with the transpose of that map
  const double* off;
};

// transient coefficients handed to the form kernels (see TanCoef / DynCoef in device_model.cu)
struct BeamDyn {
  int k_on; double at, a0, ac;         // tangent: at kv + a0 kv0 + ac kvK
  int r_on; double bK, bK0, bKc;       // resisting force: + T^T (bK kv + bK0 kv0 + bKc kvK) T v
  const double* V;                     // trial velocities [nn][ndf]
};

__device__ __forceinline__ void lobatto_rule(int n, double* xi, double* wt) {
  // LobattoBeamIntegration.cpp: the literals of getSectionLocations / getSectionWeights
  const double X[11][10] = {{0},{0},{-1.0,1.0},{-1.0,0.0,1.0},{-1.0,-0.44721360,0.44721360,1.0},
    {-1.0,-0.65465367,0.0,0.65465367,1.0},{-1.0,-0.7650553239,-0.2852315164,0.2852315164,0.7650553239,1.0},
    {-1.0,-0.8302238962,-0.4688487934,0.0,0.4688487934,0.8302238962,1.0},
    {-1.0,-0.8717401485,-0.5917001814,-0.2092992179,0.2092992179,0.5917001814,0.8717401485,1.0},
    {-1.0,-0.8997579954,-0.6771862795,-0.3631174638,0.0,0.3631174638,0.6771862795,0.8997579954,1.0},
    {-1.0,-0.9195339082,-0.7387738651,-0.4779249498,-0.1652789577,0.1652789577,0.4779249498,0.7387738651,0.9195339082,1.0}};
  const double W[11][10] = {{0},{0},{1.0,1.0},{0.333333333333333,1.333333333333333,0.333333333333333},
    {0.166666666666667,0.833333333333333,0.833333333333333,0.166666666666667},
    {0.1,0.5444444444,0.7111111111,0.5444444444,0.1},
    {0.06666666667,0.3784749562,0.5548583770,0.5548583770,0.3784749562,0.06666666667},
    {0.04761904762,0.2768260473,0.4317453812,0.4876190476,0.4317453812,0.2768260473,0.04761904762},
    {0.03571428571,0.2107042271,0.3411226924,0.4124587946,0.4124587946,0.3411226924,0.2107042271,0.03571428571},
    {0.02777777778,0.1654953615,0.2745387125,0.3464285109,0.3715192743,0.3464285109,0.2745387125,0.1654953615,0.02777777778},
    {0.02222222222,0.1333059908,0.2248893421,0.2920426836,0.3275397611,0.3275397611,0.2920426836,0.2248893421,0.1333059908,0.02222222222}};
  for (int i = 0; i < n; i++) { xi[i] = 0.5 * (X[n][i] + 1.0); wt[i] = W[n][i] * 0.5; }
}

// Steel02::setTrialStrain.  C = committed record, T = trial record (both strided by n)
// committed record of one fibre in registers (what the trial computation reads): c[0..10] as laid out above;
// t8, t9 = the previous TRIAL tangent / stress (Concrete02 only)
struct FibRec { double c[11]; double t8, t9; };
__device__ __forceinline__ void fib_load(int kind, const double* C, const double* T, long long n, FibRec& r) {
  if (kind == 0) {
#pragma unroll
    for (int v = 0; v < 8; v++) r.c[v] = C[(size_t)v * n];
    r.c[9] = C[9 * n]; r.c[10] = C[10 * n];
  } else {
    r.c[0] = C[0]; r.c[1] = C[1 * n]; r.c[9] = C[9 * n]; r.c[10] = C[10 * n];
    r.t8 = T[8 * n]; r.t9 = T[9 * n];
  }
}
__device__ __forceinline__ void steel02_trial_r(const double* __restrict__ p, const FibRec& r, double* T, long long n,
                                                double trialStrain, double& sig_o, double& e_o);
__device__ __forceinline__ void concrete02_trial_r(const double* __restrict__ p, const FibRec& r, double* T, long long n,
                                                   double trialStrain, double& sig_o, double& e_o);
__device__ __forceinline__ void steel02_trial(const double* __restrict__ p, const double* C, double* T, long long n,
                                              double trialStrain, double& sig_o, double& e_o) {
  FibRec r; fib_load(0, C, T, n, r);
  steel02_trial_r(p, r, T, n, trialStrain, sig_o, e_o);
}
__device__ __forceinline__ void steel02_trial_r(const double* __restrict__ p, const FibRec& r, double* T, long long n,
                                                double trialStrain, double& sig_o, double& e_o) {
  const double Fy = p[0], E0 = p[1], b = p[2], R0 = p[3], cR1 = p[4], cR2 = p[5], a1 = p[6], a2 = p[7], a3 = p[8],
               a4 = p[9], sigini = p[10];
  const double Esh = b * E0, epsy = Fy / E0;
  double eps = trialStrain;
  if (sigini != 0.0) eps = trialStrain + sigini / E0;
  const double epsP = r.c[10], sigP = r.c[9];
  const double deps = eps - epsP;
  double epsmin = r.c[0], epsmax = r.c[1], epspl = r.c[2], epss0 = r.c[3], sigs0 = r.c[4], epsr = r.c[5],
         sigr = r.c[6];
  int kon = (int)r.c[7];
  double sig, e;
  bool done = false;
  if (kon == 0 || kon == 3) {
    if (fabs(deps) < 10.0 * DBL_EPSILON) {
      e = E0; sig = sigini; kon = 3; done = true;
    } else {
      epsmax = epsy; epsmin = -epsy;
      if (deps < 0.0) { kon = 2; epss0 = epsmin; sigs0 = -Fy; epspl = epsmin; }
      else { kon = 1; epss0 = epsmax; sigs0 = Fy; epspl = epsmax; }
    }
  }
  if (!done) {
    if (kon == 2 && deps > 0.0) {
      kon = 1; epsr = epsP; sigr = sigP;
      if (epsP < epsmin) epsmin = epsP;
      const double d1 = (epsmax - epsmin) / (2.0 * (a4 * epsy));
      const double shft = 1.0 + a3 * pow(d1, 0.8);
      epss0 = (Fy * shft - Esh * epsy * shft - sigr + E0 * epsr) / (E0 - Esh);
      sigs0 = Fy * shft + Esh * (epss0 - epsy * shft);
      epspl = epsmax;
    } else if (kon == 1 && deps < 0.0) {
      kon = 2; epsr = epsP; sigr = sigP;
      if (epsP > epsmax) epsmax = epsP;
      const double d1 = (epsmax - epsmin) / (2.0 * (a2 * epsy));
      const double shft = 1.0 + a1 * pow(d1, 0.8);
      epss0 = (-Fy * shft + Esh * epsy * shft - sigr + E0 * epsr) / (E0 - Esh);
      sigs0 = -Fy * shft + Esh * (epss0 + epsy * shft);
      epspl = epsmin;
    }
    const double xi = fabs((epspl - epss0) / epsy);
    const double R = R0 * (1.0 - (cR1 * xi) / (cR2 + xi));
    const double epsrat = (eps - epsr) / (epss0 - epsr);
    const double dum1 = 1.0 + pow(fabs(epsrat), R);
    const double dum2 = pow(dum1, (1 / R));
    sig = b * epsrat + (1.0 - b) * epsrat / dum2;
    sig = sig * (sigs0 - sigr) + sigr;
    e = b + (1.0 - b) / (dum1 * dum2);
    e = e * (sigs0 - sigr) / (epss0 - epsr);
  }
  T[0] = epsmin; T[1 * n] = epsmax; T[2 * n] = epspl; T[3 * n] = epss0; T[4 * n] = sigs0; T[5 * n] = epsr; T[6 * n] = sigr;
  T[7 * n] = (double)kon; T[8 * n] = e; T[9 * n] = sig; T[10 * n] = eps;
  sig_o = sig; e_o = e;
}

__device__ __forceinline__ void c02_tens(const double* p, double epsc, double& sigc, double& Ect) {
  const double fc = p[0], epsc0 = p[1], ft = p[5], Ets = p[6];
  const double Ec0 = 2.0 * fc / epsc0;
  const double eps0 = ft / Ec0;
  const double epsu = ft * (1.0 / Ets + 1.0 / Ec0);
  if (epsc <= eps0) { sigc = epsc * Ec0; Ect = Ec0; }
  else if (epsc <= epsu) { Ect = -Ets; sigc = ft - Ets * (epsc - eps0); }
  else { Ect = 1.0e-10; sigc = 0.0; }
}
__device__ __forceinline__ void c02_compr(const double* p, double epsc, double& sigc, double& Ect) {
  const double fc = p[0], epsc0 = p[1], fcu = p[2], epscu = p[3];
  const double Ec0 = 2.0 * fc / epsc0;
  const double ratLocal = epsc / epsc0;
  if (epsc >= epsc0) { sigc = fc * ratLocal * (2.0 - ratLocal); Ect = Ec0 * (1.0 - ratLocal); }
  else if (epsc > epscu) { sigc = (fcu - fc) * (epsc - epsc0) / (epscu - epsc0) + fc; Ect = (fcu - fc) / (epscu - epsc0); }
  else { sigc = fcu; Ect = 1.0e-10; }
}
// Concrete02::setTrialStrain.  p = fc, epsc0, fcu, epscu (already made negative on the host), rat, ft, Ets
__device__ __forceinline__ void concrete02_trial(const double* __restrict__ p, const double* C, double* T, long long n,
                                                 double trialStrain, double& sig_o, double& e_o) {
  FibRec r; fib_load(1, C, T, n, r);
  concrete02_trial_r(p, r, T, n, trialStrain, sig_o, e_o);
}
__device__ __forceinline__ void concrete02_trial_r(const double* __restrict__ p, const FibRec& r, double* T, long long n,
                                                   double trialStrain, double& sig_o, double& e_o) {
  const double fc = p[0], epsc0 = p[1], fcu = p[2], epscu = p[3], rat = p[4];
  const double ec0 = fc * 2. / epsc0;
  double ecmin = r.c[0], dept = r.c[1];
  const double epsP = r.c[10], sigP = r.c[9];
  const double eps = trialStrain;
  const double deps = eps - epsP;
  // the early return keeps the previous TRIAL stress / tangent (Concrete02.cpp:183)
  double sig = r.t9, e = r.t8;
  if (!(fabs(deps) < DBL_EPSILON)) {
    if (eps < ecmin) {
      c02_compr(p, eps, sig, e);
      ecmin = eps;
    } else {
      const double epsr = (fcu - rat * ec0 * epscu) / (ec0 * (1.0 - rat));
      const double sigmr = ec0 * epsr;
      double sigmm, dumy;
      c02_compr(p, ecmin, sigmm, dumy);
      const double er = (sigmm - sigmr) / (ecmin - epsr);
      const double ept = ecmin - sigmm / er;
      if (eps <= ept) {
        const double sigmin = sigmm + er * (eps - ecmin);
        const double sigmax = er * .5f * (eps - ept);
        sig = sigP + ec0 * deps;
        e = ec0;
        if (sig <= sigmin) { sig = sigmin; e = er; }
        if (sig >= sigmax) { sig = sigmax; e = 0.5 * er; }
      } else {
        const double epn = ept + dept;
        double sicn;
        if (eps <= epn) {
          c02_tens(p, dept, sicn, e);
          if (dept != 0.0) e = sicn / dept; else e = ec0;
          sig = e * (eps - ept);
        } else {
          const double epstmp = eps - ept;
          c02_tens(p, epstmp, sig, e);
          dept = eps - ept;
        }
      }
    }
  }
  T[0] = ecmin; T[1 * n] = dept; T[8 * n] = e; T[9 * n] = sig; T[10 * n] = eps;
  sig_o = sig; e_o = e;
}

// Steel01::setTrialStrain + determineTrialState (Steel01.cpp:68-89, 122-196).  p = fy, E0, b, a1, a2, a3, a4
__device__ __forceinline__ void steel01_trial(const double* __restrict__ p, const double* C, double* T, long long n,
                                              double strain, double& sig_o, double& e_o) {
  const double fy = p[0], E0 = p[1], b = p[2], a1 = p[3], a2 = p[4], a3 = p[5], a4 = p[6];
  double minStrain = C[0], maxStrain = C[1 * n], shiftP = C[2 * n], shiftN = C[3 * n];
  int loading = (int)C[4 * n];
  const double Cstrain = C[10 * n], Cstress = C[9 * n];
  double Tstrain = Cstrain, Tstress = Cstress, Ttangent = C[8 * n];
  const double dStrain = strain - Cstrain;
  if (fabs(dStrain) > DBL_EPSILON) {
    Tstrain = strain;
    const double fyOneMinusB = fy * (1.0 - b);
    const double Esh = b * E0;
    const double epsy = fy / E0;
    const double c1 = Esh * Tstrain;
    const double c2 = shiftN * fyOneMinusB;
    const double c3 = shiftP * fyOneMinusB;
    const double c = Cstress + E0 * dStrain;
    const double c1c3 = c1 + c3;
    if (c1c3 < c) Tstress = c1c3; else Tstress = c;
    const double c1c2 = c1 - c2;
    if (c1c2 > Tstress) Tstress = c1c2;
    if (fabs(Tstress - c) < DBL_EPSILON) Ttangent = E0; else Ttangent = Esh;
    if (loading == 0 && dStrain != 0.0) loading = dStrain > 0.0 ? 1 : -1;
    if (loading == 1 && dStrain < 0.0) {
      loading = -1;
      if (Cstrain > maxStrain) maxStrain = Cstrain;
      shiftN = 1 + a1 * pow((maxStrain - minStrain) / (2.0 * a2 * epsy), 0.8);
    }
    if (loading == -1 && dStrain > 0.0) {
      loading = 1;
      if (Cstrain < minStrain) minStrain = Cstrain;
      shiftP = 1 + a3 * pow((maxStrain - minStrain) / (2.0 * a4 * epsy), 0.8);
    }
  }
  T[0] = minStrain; T[1 * n] = maxStrain; T[2 * n] = shiftP; T[3 * n] = shiftN; T[4 * n] = (double)loading;
  T[8 * n] = Ttangent; T[9 * n] = Tstress; T[10 * n] = Tstrain;
  sig_o = Tstress; e_o = Ttangent;
}
// ElasticMaterial (ElasticMaterial.cpp:137-182, eta = 0).  p = Epos, eta, Eneg.  A fibre section calls setTrial(), whose
// tangent at a strain of exactly zero is Epos (:146-160); a section Aggregator calls setTrialStrain() + getTangent(),
// which gives max(Epos, Eneg) there (:174-182)
__device__ __forceinline__ void elastic_trial(const double* __restrict__ p, double* T, long long n, double strain,
                                              double& sig_o, double& e_o, bool in_fibre) {
  const double Epos = p[0], Eneg = p[2];
  const double sig = strain >= 0.0 ? Epos * strain : Eneg * strain;
  const double e = strain > 0.0 ? Epos : (strain < 0.0 ? Eneg : (in_fibre ? Epos : (Epos > Eneg ? Epos : Eneg)));
  T[8 * n] = e; T[9 * n] = sig; T[10 * n] = strain;
  sig_o = sig; e_o = e;
}
// ElasticPPMaterial::setTrialStrain (ElasticPPMaterial.cpp:123-167).  p = E, fyp, fyn, ezero (host).  The plastic strain ep
// moves at commitState (:190-224) from the trial strain of that moment: the trial record carries the value the commit will
// give it (T[0]); the commit is the copy trial -> committed like every other fibre record
__device__ __forceinline__ void elasticpp_trial(const double* __restrict__ p, const double* C, double* T, long long n,
                                                double strain, double& sig_o, double& e_o) {
  const double E = p[0], fyp = p[1], fyn = p[2], ezero = p[3];
  const double ep = C[0];
  const double sigtrial = E * (strain - ezero - ep);
  const double f = sigtrial >= 0.0 ? sigtrial - fyp : -sigtrial + fyn;
  const double fYieldSurface = -E * DBL_EPSILON;
  double sig, e, epn = ep;
  if (f <= fYieldSurface) { sig = sigtrial; e = E; }
  else {
    sig = sigtrial > 0.0 ? fyp : fyn; e = 0.0;
    if (sigtrial > 0.0) epn += f / E; else epn -= f / E;
  }
  T[0] = epn; T[8 * n] = e; T[9 * n] = sig; T[10 * n] = strain;
  sig_o = sig; e_o = e;
}
// Concrete01::setTrialStrain with reload / envelope / unload (Concrete01.cpp:146-206, 313-385).  p = fpc, epsc0, fpcu, epscu
// (made negative on the host)
__device__ __forceinline__ void concrete01_trial(const double* __restrict__ p, const double* C, double* T, long long n,
                                                 double strain, double& sig_o, double& e_o) {
  const double fpc = p[0], epsc0 = p[1], fpcu = p[2], epscu = p[3];
  const double CminStrain = C[0], CendStrain = C[1 * n], CunloadSlope = C[2 * n];
  const double Cstrain = C[10 * n], Cstress = C[9 * n];
  double TminStrain = CminStrain, TendStrain = CendStrain, TunloadSlope = CunloadSlope;
  double Tstress = Cstress, Ttangent = C[8 * n], Tstrain = Cstrain;
  const double dStrain = strain - Cstrain;
  if (!(fabs(dStrain) < DBL_EPSILON)) {
    Tstrain = strain;
    if (Tstrain > 0.0) { Tstress = 0; Ttangent = 0; }
    else {
      const double tempStress = Cstress + TunloadSlope * Tstrain - TunloadSlope * Cstrain;
      if (strain < Cstrain) {
        // reload()
        if (Tstrain <= TminStrain) {
          TminStrain = Tstrain;
          // envelope()
          if (Tstrain > epsc0) {
            const double eta = Tstrain / epsc0;
            Tstress = fpc * (2 * eta - eta * eta);
            const double Ec0 = 2.0 * fpc / epsc0;
            Ttangent = Ec0 * (1.0 - eta);
          } else if (Tstrain > epscu) {
            Ttangent = (fpc - fpcu) / (epsc0 - epscu);
            Tstress = fpc + Ttangent * (Tstrain - epsc0);
          } else { Tstress = fpcu; Ttangent = 0.0; }
          // unload()
          double tempStrain = TminStrain;
          if (tempStrain < epscu) tempStrain = epscu;
          const double eta = tempStrain / epsc0;
          double ratio = 0.707 * (eta - 2.0) + 0.834;
          if (eta < 2.0) ratio = 0.145 * eta * eta + 0.13 * eta;
          TendStrain = ratio * epsc0;
          const double temp1 = TminStrain - TendStrain;
          const double Ec0 = 2.0 * fpc / epsc0;
          const double temp2 = Tstress / Ec0;
          if (temp1 > -DBL_EPSILON) TunloadSlope = Ec0;
          else if (temp1 <= temp2) { TendStrain = TminStrain - temp1; TunloadSlope = Tstress / temp1; }
          else { TendStrain = TminStrain - temp2; TunloadSlope = Ec0; }
        } else if (Tstrain <= TendStrain) {
          Ttangent = TunloadSlope;
          Tstress = Ttangent * (Tstrain - TendStrain);
        } else { Tstress = 0.0; Ttangent = 0.0; }
        if (tempStress > Tstress) { Tstress = tempStress; Ttangent = TunloadSlope; }
      } else if (tempStress <= 0.0) { Tstress = tempStress; Ttangent = TunloadSlope; }
      else { Tstress = 0.0; Ttangent = 0.0; }
    }
  }
  T[0] = TminStrain; T[1 * n] = TendStrain; T[2 * n] = TunloadSlope;
  T[8 * n] = Ttangent; T[9 * n] = Tstress; T[10 * n] = Tstrain;
  sig_o = Tstress; e_o = Ttangent;
}
__device__ __forceinline__ void uniaxial_trial(int kind, const double* __restrict__ p, const double* C, double* T, long long n,
                                               double strain, double& stress, double& tangent, bool in_fibre = true) {
  if (kind == 0) steel02_trial(p, C, T, n, strain, stress, tangent);
  else if (kind == 1) concrete02_trial(p, C, T, n, strain, stress, tangent);
  else if (kind == 2) steel01_trial(p, C, T, n, strain, stress, tangent);
  else if (kind == 4) concrete01_trial(p, C, T, n, strain, stress, tangent);
  else if (kind == 5) elasticpp_trial(p, C, T, n, strain, stress, tangent);
  else elastic_trial(p, T, n, strain, stress, tangent, in_fibre);
}

// displacements (or velocities) of the element ends from those of the nodes: u_end = u + theta x offset (2D)
__device__ __forceinline__ void fbc2d_end_disp(const BeamView& B, long long e, double* ug) {
  if (!B.off) return;
  const double oix = B.off[e], oiy = B.off[B.n + e], ojx = B.off[2 * B.n + e], ojy = B.off[3 * B.n + e];
  ug[0] += -ug[2] * oiy; ug[1] += ug[2] * oix;
  ug[3] += -ug[5] * ojy; ug[4] += ug[5] * ojx;
}

// the same in 3D: u_end = u + theta x offset, per node
__device__ __forceinline__ void fbc3d_end_disp(const BeamView& B, long long e, double* ug) {
  if (!B.off) return;
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const double dx = B.off[(size_t)(3 * a) * B.n + e], dy = B.off[(size_t)(3 * a + 1) * B.n + e], dz = B.off[(size_t)(3 * a + 2) * B.n + e];
    double* u = ug + 6 * a;
    const double tx = u[3], ty = u[4], tz = u[5];
    u[0] += ty * dz - tz * dy; u[1] += tz * dx - tx * dz; u[2] += tx * dy - ty * dx;
  }
}

// FiberSection2d::setTrialSectionDeformation for section i of element e -> s[2], k[4] (column-major)
__device__ __forceinline__ void section_trial(const BeamView& B, long long e, int i, const double* d, double* s, double* k) {
  k[0] = k[1] = k[2] = k[3] = 0.0; s[0] = s[1] = 0.0;
  const double d0 = d[0], d1 = d[1];
  for (int f = 0; f < B.nf; f++) {
    const double y = __ldg(B.fy + f), A = __ldg(B.fA + f);
    const int kind = __ldg(B.fkind + f);
    const size_t rec = ((size_t)(i * B.nf + f) * XB_FIB_NV) * B.n + e;
    double stress, tangent;
    if (B.agg) {
      // section Aggregator (SectionAggregator.cpp:316, :365, :483): material 0 on the axial strain, material 1 on the
      // curvature; tangent and stress resultant are the materials' own, no coupling
      uniaxial_trial(kind & 15, B.fpar + f * 12, B.fc + rec, B.ft + rec, B.n, (kind & XB_FIB_CURV) ? d1 : d0, stress, tangent, false);
      if (kind & XB_FIB_CURV) { k[3] = tangent; s[1] = stress; } else { k[0] = tangent; s[0] = stress; }
      continue;
    }
    const double strain = d0 - y * d1;
    uniaxial_trial(kind, B.fpar + f * 12, B.fc + rec, B.ft + rec, B.n, strain, stress, tangent);
    const double ks0 = tangent * A;
    const double ks1 = ks0 * -y;
    k[0] += ks0; k[1] += ks1; k[3] += ks1 * -y;
    const double fs0 = stress * A;
    s[0] += fs0; s[1] += fs0 * -y;
  }
  k[2] = k[1];
}
// the section's getSectionFlexibility: inverse of the tangent, or SectionAggregator's own (SectionAggregator.cpp:419-450)
__device__ __forceinline__ void inv2(const double* a, double* ainv);
__device__ __forceinline__ void section_flex(const BeamView& B, const double* k, double* f) {
  if (B.agg) {
    f[1] = f[2] = 0.0;
    f[0] = k[0] == 0.0 ? 1.e14 : 1 / k[0];
    f[3] = k[3] == 0.0 ? 1.e14 : 1 / k[3];
    return;
  }
  inv2(k, f);
}
__device__ __forceinline__ void inv2(const double* a, double* ainv) {   // matrix/routines/invGL2.c
  const double det = a[0] * a[3] - a[2] * a[1];
  ainv[0] = a[3] / det; ainv[1] = -a[1] / det; ainv[2] = -a[2] / det; ainv[3] = a[0] / det;
}
__device__ __forceinline__ void inv3(const double* a, double* ainv) {   // matrix/routines/invGL3.c
  const double* A = a - 4; double* I = ainv - 4;
  const double det = A[4]*A[8]*A[12] - A[4]*A[11]*A[9] - A[7]*A[5]*A[12] + A[7]*A[11]*A[6] + A[10]*A[5]*A[9] - A[10]*A[8]*A[6];
  double c[9];
  c[0] =  A[8]*A[12] - A[11]*A[9];  c[3] = -(A[5]*A[12] - A[11]*A[6]); c[6] =  A[5]*A[9] - A[8]*A[6];
  c[1] = -(A[7]*A[12] - A[10]*A[9]); c[4] =  A[4]*A[12] - A[10]*A[6];  c[7] = -(A[4]*A[9] - A[7]*A[6]);
  c[2] =  A[7]*A[11] - A[10]*A[8];  c[5] = -(A[4]*A[11] - A[10]*A[5]); c[8] =  A[4]*A[8] - A[7]*A[5];
  for (int i = 1; i <= 3; ++i) for (int j = 1; j <= 3; ++j) I[j + i * 3] = c[i + j * 3 - 4] / det;
}
__device__ __forceinline__ void crd2d_basic(double L, double cosT, double sinT, const double* ug, double* ub) {
  const double oneOverL = 1.0 / L;
  const double sl = sinT * oneOverL, cl = cosT * oneOverL;
  ub[0] = -cosT * ug[0] - sinT * ug[1] + cosT * ug[3] + sinT * ug[4];
  ub[1] = -sl * ug[0] + cl * ug[1] + ug[2] + sl * ug[3] - cl * ug[4];
  ub[2] = ub[1] + ug[5] - ug[2];
}

// CorotCrdTransf2d::update (CorotCrdTransf2d.cpp:179-231, no offsets): local end displacements, the deformed chord
// (compElemtLengthAndOrientWRTLocalSystem, :272-299), basic deformations with the chord's rigid rotation taken out
// (transfLocalDisplsToBasic, :344-355).  cg = Ln, cosAlpha, sinAlpha; false: zero deformed length
__device__ __forceinline__ bool corot2d_geom(double L, double cosT, double sinT, const double* ug, double* cg, double* ub) {
  const double ul0 = cosT * ug[0] + sinT * ug[1], ul1 = cosT * ug[1] - sinT * ug[0];
  const double ul3 = cosT * ug[3] + sinT * ug[4], ul4 = cosT * ug[4] - sinT * ug[3];
  const double Lx = L + (ul3 - ul0), Ly = ul4 - ul1;
  const double Ln = sqrt(Lx * Lx + Ly * Ly);
  if (Ln == 0.0) return false;
  const double cosA = Lx / Ln, sinA = Ly / Ln;
  cg[0] = Ln; cg[1] = cosA; cg[2] = sinA;
  if (ub) {
    const double alpha = atan2(sinA, cosA);
    ub[0] = Ln - L; ub[1] = ug[2] - alpha; ub[2] = ug[5] - alpha;
  }
  return true;
}

// getTangentStiff -> LinearCrdTransf2d::getGlobalStiffMatrix(kv); getResistingForce ->
// getGlobalResistingForce(Se).  Rows of node a go to that node's slot (node-major storage).
__global__ void __launch_bounds__(128) fbc2d_form_kernel(BeamView B, int want_k, int want_r, int transpose, BeamDyn dy) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n) return;
  const long long n = B.n;
  const double L = B.geo[e], cosTheta = B.geo[n + e], sinTheta = B.geo[2 * n + e], oneOverL = 1.0 / L;
  // geomTransf Corotational: the 2D element refreshes the transformation from the nodes' trial displacements before it
  // forms its tangent or its forces (ForceBeamColumn2d.cpp:402,526)
  double Ln = L, cosA = 1.0, sinA = 0.0;
  if (B.corot) {
    double ue[6], cg[3];
    for (int j = 0; j < 3; j++) { ue[j] = B.U[(size_t)B.conn[e * 2] * 3 + j]; ue[3 + j] = B.U[(size_t)B.conn[e * 2 + 1] * 3 + j]; }
    if (corot2d_geom(L, cosTheta, sinTheta, ue, cg, nullptr)) { Ln = cg[0]; cosA = cg[1]; sinA = cg[2]; }
  }
  if (want_k && B.corot) {
    // CorotCrdTransf2d::getGlobalStiffMatrix (:525-644): kl = Tbl' kb Tbl + geometric stiffness (getGeomStiffMatrix, :908-953),
    // then local -> global.  Transient without element damping: the whole tangent times c1 (host: no stiffness-proportional
    // Rayleigh terms on corotational beams)
    const double at = dy.k_on ? dy.at : 1.0;
    double kb[9];
    for (int i = 0; i < 9; i++) kb[i] = at * B.kv[i * n + e];
    const double Tbl[3][6] = {{-cosA, -sinA, 0.0, cosA, sinA, 0.0},
                              {-sinA / Ln, cosA / Ln, 1.0, sinA / Ln, -cosA / Ln, 0.0},
                              {-sinA / Ln, cosA / Ln, 0.0, sinA / Ln, -cosA / Ln, 1.0}};
    double tk[3][6], kl[6][6];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 6; j++) tk[i][j] = kb[i] * Tbl[0][j] + kb[i + 3] * Tbl[1][j] + kb[i + 6] * Tbl[2][j];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 6; j++) kl[i][j] = Tbl[0][i] * tk[0][j] + Tbl[1][i] * tk[1][j] + Tbl[2][i] * tk[2][j];
    {
      const double s2 = sinA * sinA, c2 = cosA * cosA, cs = sinA * cosA;
      const double f0 = at * B.Se[e] / Ln, f12 = at * (B.Se[n + e] + B.Se[2 * n + e]) / (Ln * Ln);
      // the 2x2 block g on the translational dofs of a node: +g on (I,I), (J,J), -g on (I,J), (J,I)
      const double g[2][2] = {{s2 * f0 - 2 * cs * f12, -cs * f0 + (c2 - s2) * f12}, {-cs * f0 + (c2 - s2) * f12, c2 * f0 + 2 * cs * f12}};
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { kl[i][j] += g[i][j]; kl[3 + i][3 + j] += g[i][j]; kl[i][3 + j] -= g[i][j]; kl[3 + i][j] -= g[i][j]; }
    }
    double kg[6][6];
    const double S2 = sinTheta * sinTheta, C2 = cosTheta * cosTheta, CS = sinTheta * cosTheta;
#pragma unroll
    for (int bi = 0; bi < 2; bi++)
#pragma unroll
      for (int bj = 0; bj < 2; bj++) {
        const int r = 3 * bi, c = 3 * bj;
        const double k11 = kl[r][c], k12 = kl[r][c + 1], k13 = kl[r][c + 2], k21 = kl[r + 1][c], k22 = kl[r + 1][c + 1], k23 = kl[r + 1][c + 2],
                     k31 = kl[r + 2][c], k32 = kl[r + 2][c + 1], k33 = kl[r + 2][c + 2];
        kg[r][c] = C2 * k11 + S2 * k22 - CS * (k21 + k12);
        kg[r + 1][c] = C2 * k21 - S2 * k12 + CS * (k11 - k22);
        kg[r + 2][c] = cosTheta * k31 - sinTheta * k32;
        kg[r][c + 1] = C2 * k12 - S2 * k21 + CS * (k11 - k22);
        kg[r + 1][c + 1] = C2 * k22 + S2 * k11 + CS * (k21 + k12);
        kg[r + 2][c + 1] = sinTheta * k31 + cosTheta * k32;
        kg[r][c + 2] = cosTheta * k13 - sinTheta * k23;
        kg[r + 1][c + 2] = sinTheta * k13 + cosTheta * k23;
        kg[r + 2][c + 2] = k33;
      }
    for (int a = 0; a < 2; a++) {
      const long long d = B.kdst[e * 2 + a];
      double* base = d >= 0 ? B.KeN + d : B.sendK + (-d - 1);
      for (int p = 0; p < 3; p++)
        for (int c = 0; c < 6; c++) base[p * B.cps + c] = transpose ? kg[c][a * 3 + p] : kg[a * 3 + p][c];
    }
  } else if (want_k) {
    double kb[9];
    for (int i = 0; i < 9; i++) kb[i] = B.kv[i * n + e];
    if (dy.k_on) {
      for (int i = 0; i < 9; i++) {
        double t = dy.at * kb[i];
        if (dy.a0 != 0.0) t += dy.a0 * B.kv0[i * n + e];
        if (dy.ac != 0.0) t += dy.ac * B.kvK[i * n + e];
        kb[i] = t;
      }
    }
    const double kb00 = kb[0], kb10 = kb[1], kb20 = kb[2], kb01 = kb[3], kb11 = kb[4], kb21 = kb[5], kb02 = kb[6], kb12 = kb[7], kb22 = kb[8];
    double tmp[3][6], kg[6][6];
    const double sl = sinTheta * oneOverL, cl = cosTheta * oneOverL;
    tmp[0][0] = -cosTheta * kb00 - sl * (kb01 + kb02); tmp[0][1] = -sinTheta * kb00 + cl * (kb01 + kb02);
    tmp[0][2] = kb01; tmp[0][3] = -tmp[0][0]; tmp[0][4] = -tmp[0][1]; tmp[0][5] = kb02;
    tmp[1][0] = -cosTheta * kb10 - sl * (kb11 + kb12); tmp[1][1] = -sinTheta * kb10 + cl * (kb11 + kb12);
    tmp[1][2] = kb11; tmp[1][3] = -tmp[1][0]; tmp[1][4] = -tmp[1][1]; tmp[1][5] = kb12;
    tmp[2][0] = -cosTheta * kb20 - sl * (kb21 + kb22); tmp[2][1] = -sinTheta * kb20 + cl * (kb21 + kb22);
    tmp[2][2] = kb21; tmp[2][3] = -tmp[2][0]; tmp[2][4] = -tmp[2][1]; tmp[2][5] = kb22;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      kg[0][c] = -cosTheta * tmp[0][c] - sl * (tmp[1][c] + tmp[2][c]);
      kg[1][c] = -sinTheta * tmp[0][c] + cl * (tmp[1][c] + tmp[2][c]);
      kg[2][c] = tmp[1][c];
      kg[3][c] = -kg[0][c]; kg[4][c] = -kg[1][c];
      kg[5][c] = tmp[2][c];
    }
    if (B.pdelta) {
      // PDeltaCrdTransf2d::getGlobalStiffMatrix (PDeltaCrdTransf2d.cpp:628-633): N/L on the local transverse dofs,
      // kl[1][1], kl[4][4] += N/L, kl[1][4], kl[4][1] -= N/L; in global axes N/L t t' with t = (-sin, cos)
      // (transient with element damping: (c1 + c2 betaK) Kt + c2 betaKc Kc carry their own axial forces; the initial
      //  stiffness has no geometric part, PDeltaCrdTransf2d::getInitialGlobalStiffMatrix)
      const double NoverL = (dy.k_on ? dy.at * B.Se[e] + (dy.ac != 0.0 ? dy.ac * B.nK[e] : 0.0) : B.Se[e]) * oneOverL;
      const double t[2] = {-sinTheta, cosTheta};
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const double g = NoverL * t[i] * t[j];
          kg[i][j] += g; kg[3 + i][3 + j] += g; kg[i][3 + j] -= g; kg[3 + i][j] -= g;
        }
    }
    if (B.off) {   // K_node = To' K_end To (the t02, t12, t35, t45 terms of LinearCrdTransf2d::getGlobalStiffMatrix)
      const double cf[2][2] = {{-B.off[n + e], B.off[e]}, {-B.off[3 * n + e], B.off[2 * n + e]}};
      for (int i = 0; i < 6; i++) for (int a = 0; a < 2; a++) kg[i][3 * a + 2] += kg[i][3 * a] * cf[a][0] + kg[i][3 * a + 1] * cf[a][1];
      for (int j = 0; j < 6; j++) for (int a = 0; a < 2; a++) kg[3 * a + 2][j] += cf[a][0] * kg[3 * a][j] + cf[a][1] * kg[3 * a + 1][j];
    }
    for (int a = 0; a < 2; a++) {
      const long long d = B.kdst[e * 2 + a];
      double* base = d >= 0 ? B.KeN + d : B.sendK + (-d - 1);
      for (int p = 0; p < 3; p++)
        for (int c = 0; c < 6; c++) base[p * B.cps + c] = transpose ? kg[c][a * 3 + p] : kg[a * 3 + p][c];
    }
  }
  if (want_r) {
    double q0 = B.Se[e], q1 = B.Se[n + e], q2 = B.Se[2 * n + e];
    const double q0s = q0;               // the element's own axial force (the P-Delta shear is not a damping term)
    double gdv = 0.0;                    // PDelta: geometric part of (bK Kt + bKc Kc) v, on the local transverse dofs
    if (dy.r_on) {
      // Element::getRayleighDampingForces with stiffness-proportional terms: T^T [kd (T v)], kd basic
      double vg[6], vb[3];
      for (int a = 0; a < 2; a++) {
        const int nd = B.conn[e * 2 + a];
        for (int j = 0; j < 3; j++) vg[a * 3 + j] = dy.V[(size_t)nd * 3 + j];
      }
      fbc2d_end_disp(B, e, vg);
      crd2d_basic(L, cosTheta, sinTheta, vg, vb);
      if (B.pdelta) {
        const double gd = (dy.bK * B.Se[e] + (dy.bKc != 0.0 ? dy.bKc * B.nK[e] : 0.0)) * oneOverL;
        gdv = gd * ((-sinTheta * vg[0] + cosTheta * vg[1]) - (-sinTheta * vg[3] + cosTheta * vg[4]));
      }
      double qd[3] = {0, 0, 0};
      for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) {
          double kd = dy.bK * B.kv[(r + 3 * c) * n + e];
          if (dy.bK0 != 0.0) kd += dy.bK0 * B.kv0[(r + 3 * c) * n + e];
          if (dy.bKc != 0.0) kd += dy.bKc * B.kvK[(r + 3 * c) * n + e];
          qd[r] += kd * vb[c];
        }
      q0 += qd[0]; q1 += qd[1]; q2 += qd[2];
    }
    const double V = oneOverL * (q1 + q2);
    double pl[6] = {-q0, V, q1, q0, -V, q2};
    if (B.corot) {   // CorotCrdTransf2d::getGlobalResistingForce (:483-522): pl = Tbl' pb on the deformed chord
      const double Vc = (q1 + q2) / Ln;
      pl[0] = -cosA * q0 - sinA * Vc; pl[1] = -sinA * q0 + cosA * Vc; pl[2] = q1;
      pl[3] = cosA * q0 + sinA * Vc; pl[4] = sinA * q0 - cosA * Vc; pl[5] = q2;
    }
    if (B.wl != nullptr && B.loads_on) {   // computeReactions (ForceBeamColumn2d.cpp:407-425) into LinearCrdTransf2d's pl
      const double wy = B.wl[e] * B.lam, wa = B.wl[2 * n + e] * B.lam;
      double p0[3] = {0.0, 0.0, 0.0};
      p0[0] -= wa * L;
      const double Vr = 0.5 * wy * L;
      p0[1] -= Vr; p0[2] -= Vr;
      if (B.has_point) {   // Beam2dPointLoad (ForceBeamColumn2d.cpp:442-455); no point load on this element: zeros
        const double P = B.wl[3 * n + e] * B.lam, N = B.wl[5 * n + e] * B.lam, aOverL = B.wl[6 * n + e];
        const double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
        p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
      }
      if (B.has_partial && B.wl[12 * n + e] > B.wl[11 * n + e]) {   // Beam2dPartialUniformLoad (ForceBeamColumn2d.cpp:426-443)
        const double wya = B.wl[7 * n + e] * B.lam, wyb = B.wl[8 * n + e] * B.lam, waa = B.wl[9 * n + e] * B.lam, wab = B.wl[10 * n + e] * B.lam;
        const double a = B.wl[11 * n + e] * L, b = B.wl[12 * n + e] * L;
        p0[0] -= waa * (b - a) + 0.5 * (wab - waa) * (b - a);
        double Fy = wya * (b - a);
        double c = a + 0.5 * (b - a);
        p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
        Fy = 0.5 * (wyb - wya) * (b - a);
        c = a + 2.0 / 3.0 * (b - a);
        p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
      }
      pl[0] += p0[0]; pl[1] += p0[1]; pl[4] += p0[2];
    }
    if (B.pdelta) {   // PDeltaCrdTransf2d::update + getGlobalResistingForce (PDeltaCrdTransf2d.cpp:349-384, 532-535)
      double ue[6];
      for (int j = 0; j < 3; j++) { ue[j] = B.U[(size_t)B.conn[e * 2] * 3 + j]; ue[3 + j] = B.U[(size_t)B.conn[e * 2 + 1] * 3 + j]; }
      fbc2d_end_disp(B, e, ue);
      const double* uI = ue; const double* uJ = ue + 3;
      const double ul1 = -sinTheta * uI[0] + cosTheta * uI[1];
      const double ul4 = -sinTheta * uJ[0] + cosTheta * uJ[1];
      const double NoverL = (ul1 - ul4) * q0s * oneOverL;
      pl[1] += NoverL; pl[4] -= NoverL;
      pl[1] += gdv; pl[4] -= gdv;
    }
    double* R = B.Re + e * 6;
    R[0] = cosTheta * pl[0] - sinTheta * pl[1];
    R[1] = sinTheta * pl[0] + cosTheta * pl[1];
    R[2] = pl[2];
    R[3] = cosTheta * pl[3] - sinTheta * pl[4];
    R[4] = sinTheta * pl[3] + cosTheta * pl[4];
    R[5] = pl[5];
    if (B.off) {   // R_node = To' R_end: the end forces' moments about the nodes
      R[2] += -B.off[n + e] * R[0] + B.off[e] * R[1];
      R[5] += -B.off[3 * n + e] * R[3] + B.off[2 * n + e] * R[4];
    }
  }
}

// ForceBeamColumn2d::revertToLastCommit: the fibre records have been copied back already
// (committed -> trial); sections recompute (FiberSection2d::revertToLastCommit, then
// setTrialSectionDeformation(vs)), element state back, initialFlag = 0.
__global__ void __launch_bounds__(64) fbc2d_revert_kernel(BeamView B) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n) return;
  const long long n = B.n;
  for (int i = 0; i < B.nip; i++) {
    double vs[2], s[2], k[4], fl[4];
    for (int q = 0; q < 2; q++) { vs[q] = B.vsc[(size_t)(i * 2 + q) * n + e]; B.vs[(size_t)(i * 2 + q) * n + e] = vs[q]; }
    section_trial(B, e, i, vs, s, k);
    section_flex(B, k, fl);
    for (int q = 0; q < 2; q++) B.Ssr[(size_t)(i * 2 + q) * n + e] = s[q];
    for (int q = 0; q < 4; q++) B.fs[(size_t)(i * 4 + q) * n + e] = fl[q];
  }
  for (int i = 0; i < 3; i++) B.Se[i * n + e] = B.Sec[i * n + e];
  for (int i = 0; i < 9; i++) B.kv[i * n + e] = B.kvc[i * n + e];
  B.iflag[e] = 0;
}

// =====================================================================================
// 3D: ForceBeamColumn3d + FiberSection3d (P, Mz, My, T with elastic torsion) + LinearCrdTransf3d
//   ForceBeamColumn3d::update / commitState / revertToLastCommit / getTangentStiff / getResistingForce
//                                         element/Frame/Other/Force/ForceBeamColumn3d.cpp:587,279,313,401,551
//   LinearCrdTransf3d (no offsets)        coordTransformation/LinearCrdTransf3d.cpp:344,492,699,767
//   FiberSection3d::setTrialSectionDeformation      material/section/FiberSection3d.cpp:422
//   Matrix::Invert -> cmx_inv4 / cmx_inv6 (cofactor expansions, matrix/routines/invGL4.c, invGL6.c): the
//   section stiffness and the element flexibility are block diagonal here (torsion uncoupled), so the
//   P-Mz-My block goes through the 3x3 cofactor formula and the 5x5 block of the flexibility through
//   Gauss-Jordan elimination with partial pivoting; agreement is to rounding x condition number.
// =====================================================================================

// FiberSection3d::setTrialSectionDeformation for section i of element e -> s[4], k[16] (column-major).
// (Requesting fibre f+1's record ahead of fibre f's update was tried: the extra live registers cost more
// than the hidden latency gains, 3.5 vs 3.0 ms on the 195k-element frame.)
__device__ __forceinline__ void section3_trial(const BeamView& B, long long e, int i, const double* d, double* s, double* k) {
  for (int q = 0; q < 16; q++) k[q] = 0.0;
  s[0] = s[1] = s[2] = 0.0;
  const double e0 = d[0], e1 = d[1], e2 = d[2], e3 = d[3];
  for (int f = 0; f < B.nf; f++) {
    const double y = __ldg(B.fy + f), z = __ldg(B.fz + f), A = __ldg(B.fA + f);
    const double strain = e0 - y * e1 + z * e2;
    const size_t rec = ((size_t)(i * B.nf + f) * XB_FIB_NV) * B.n + e;
    double stress, tangent;
    uniaxial_trial(__ldg(B.fkind + f), B.fpar + f * 12, B.fc + rec, B.ft + rec, B.n, strain, stress, tangent);
    const double EA = tangent * A;
    k[0] += EA; k[1] += -y * EA; k[2] += z * EA;
    k[5] += y * y * EA; k[10] += z * z * EA; k[6] += -y * z * EA;
    const double fs0 = stress * A;
    s[0] += fs0; s[1] += -y * fs0; s[2] += z * fs0;
  }
  k[4] = k[1]; k[8] = k[2]; k[9] = k[6];
  s[3] = B.GJ * e3; k[15] = B.GJ;
}
__device__ __forceinline__ void section3_flex(const double* k, double* f) {
  double a[9], ai[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) a[r + 3 * c] = k[r + 4 * c];
  inv3(a, ai);
  for (int q = 0; q < 16; q++) f[q] = 0.0;
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) f[r + 4 * c] = ai[r + 3 * c];
  f[15] = 1.0 / k[15];
}
// inverse of the element flexibility (column-major 6x6, torsion uncoupled); false if singular
__device__ __forceinline__ bool inv6_flex(const double* f, double* kv) {
  double a[5][10];
  for (int r = 0; r < 5; r++) for (int c = 0; c < 5; c++) { a[r][c] = f[r + 6 * c]; a[r][5 + c] = (r == c) ? 1.0 : 0.0; }
  for (int c = 0; c < 5; c++) {
    int p = c; double big = fabs(a[c][c]);
    for (int r = c + 1; r < 5; r++) if (fabs(a[r][c]) > big) { big = fabs(a[r][c]); p = r; }
    if (big == 0.0) return false;
    if (p != c) for (int q = 0; q < 10; q++) { const double t = a[c][q]; a[c][q] = a[p][q]; a[p][q] = t; }
    const double piv = 1.0 / a[c][c];
    for (int q = 0; q < 10; q++) a[c][q] *= piv;
    for (int r = 0; r < 5; r++) if (r != c) { const double mlt = a[r][c]; if (mlt != 0.0) for (int q = 0; q < 10; q++) a[r][q] -= mlt * a[c][q]; }
  }
  for (int q = 0; q < 36; q++) kv[q] = 0.0;
  for (int r = 0; r < 5; r++) for (int c = 0; c < 5; c++) kv[r + 6 * c] = a[r][5 + c];
  kv[35] = 1.0 / f[35];
  return true;
}
__device__ __forceinline__ void crd3d_basic(double L, const double* R, const double* ug, double* ub) {
  double ul[12];
  for (int blk = 0; blk < 4; blk++)
    for (int r = 0; r < 3; r++)
      ul[3 * blk + r] = R[3 * r] * ug[3 * blk] + R[3 * r + 1] * ug[3 * blk + 1] + R[3 * r + 2] * ug[3 * blk + 2];
  const double oneOverL = 1.0 / L;
  double tmp;
  ub[0] = ul[6] - ul[0];
  tmp = oneOverL * (ul[1] - ul[7]);
  ub[1] = ul[5] + tmp; ub[2] = ul[11] + tmp;
  tmp = oneOverL * (ul[8] - ul[2]);
  ub[3] = ul[4] + tmp; ub[4] = ul[10] + tmp;
  ub[5] = ul[9] - ul[3];
}
__device__ __forceinline__ double norm6(const double* v) { double s = 0.0; for (int i = 0; i < 6; i++) s += v[i] * v[i]; return sqrt(s); }

// ---- the same update with one LANE PER SECTION (G lanes per element, G = 4 | 8 | 16 >= nIP) ----
// The thread-per-element form above keeps every section's state in per-thread arrays (4 KB of local
// memory, 255 registers, one resident warp per scheduler) and walks nIP x nf fibres serially.  Here
// lane i of an element's group owns section i: its deformation, flexibility and resisting force are
// scalars in registers, the fibre loops of the sections run side by side, and the element-level
// algebra (flexibility sum, 6x6 inverse, energy test) is done redundantly by the G lanes on values
// summed over the group IN SECTION ORDER (shuffle from lane 0, 1, ..: the reference's f += .., vr += ..),
// so the arithmetic -- and every convergence decision -- is that of the thread-per-element form.
__constant__ double LOBATTO_X[11][10] = {{0},{0},{-1.0,1.0},{-1.0,0.0,1.0},{-1.0,-0.44721360,0.44721360,1.0},
    {-1.0,-0.65465367,0.0,0.65465367,1.0},{-1.0,-0.7650553239,-0.2852315164,0.2852315164,0.7650553239,1.0},
    {-1.0,-0.8302238962,-0.4688487934,0.0,0.4688487934,0.8302238962,1.0},
    {-1.0,-0.8717401485,-0.5917001814,-0.2092992179,0.2092992179,0.5917001814,0.8717401485,1.0},
    {-1.0,-0.8997579954,-0.6771862795,-0.3631174638,0.0,0.3631174638,0.6771862795,0.8997579954,1.0},
    {-1.0,-0.9195339082,-0.7387738651,-0.4779249498,-0.1652789577,0.1652789577,0.4779249498,0.7387738651,0.9195339082,1.0}};
__constant__ double LOBATTO_W[11][10] = {{0},{0},{1.0,1.0},{0.333333333333333,1.333333333333333,0.333333333333333},
    {0.166666666666667,0.833333333333333,0.833333333333333,0.166666666666667},
    {0.1,0.5444444444,0.7111111111,0.5444444444,0.1},
    {0.06666666667,0.3784749562,0.5548583770,0.5548583770,0.3784749562,0.06666666667},
    {0.04761904762,0.2768260473,0.4317453812,0.4876190476,0.4317453812,0.2768260473,0.04761904762},
    {0.03571428571,0.2107042271,0.3411226924,0.4124587946,0.4124587946,0.3411226924,0.2107042271,0.03571428571},
    {0.02777777778,0.1654953615,0.2745387125,0.3464285109,0.3715192743,0.3464285109,0.2745387125,0.1654953615,0.02777777778},
    {0.02222222222,0.1333059908,0.2248893421,0.2920426836,0.3275397611,0.3275397611,0.2920426836,0.2248893421,0.1333059908,0.02222222222}};
__device__ __forceinline__ double lobatto_x(int n, int i) { return 0.5 * (LOBATTO_X[n][i] + 1.0); }
__device__ __forceinline__ double lobatto_w(int n, int i) { return LOBATTO_W[n][i] * 0.5; }
// section i of element e: location and weight as fractions of L (BeamIntegration::getSectionLocations / getSectionWeights)
__device__ __forceinline__ double rule_x(const BeamView& B, long long e, int i) { return B.rule ? B.rule[(size_t)i * B.n + e] : lobatto_x(B.nip, i); }
__device__ __forceinline__ double rule_w(const BeamView& B, long long e, int i) { return B.rule ? B.rule[(size_t)(B.nip + i) * B.n + e] : lobatto_w(B.nip, i); }
// sum over the sections of an element in section order: lanes gbase .. gbase + nip - 1 of the warp
__device__ __forceinline__ double group_sum_ordered(double x, unsigned gmask, int gbase, int nip) {
  double s = __shfl_sync(gmask, x, gbase);
  for (int k = 1; k < nip; k++) s += __shfl_sync(gmask, x, gbase + k);
  return s;
}
// inv6_flex without dynamically indexed rows (everything stays in registers): f5 is the 5x5 block
// (column-major, stride 5), f55 the torsion entry; same pivot choice and operation order as inv6_flex
template <int C>
__device__ __forceinline__ void gj5_step(double (&a)[5][10], bool& ok) {
  // column C of the Gauss-Jordan elimination with partial pivoting; all indices are compile-time
  int p = C; double big = fabs(a[C][C]);
#pragma unroll
  for (int r = C + 1; r < 5; r++) if (fabs(a[r][C]) > big) { big = fabs(a[r][C]); p = r; }
  if (big == 0.0) ok = false;
#pragma unroll
  for (int r = C + 1; r < 5; r++) {
    const bool sw = (p == r);
#pragma unroll
    for (int q = 0; q < 10; q++) { const double x = a[C][q], y = a[r][q]; a[C][q] = sw ? y : x; a[r][q] = sw ? x : y; }
  }
  const double piv = 1.0 / a[C][C];
#pragma unroll
  for (int q = 0; q < 10; q++) a[C][q] *= piv;
#pragma unroll
  for (int r = 0; r < 5; r++)
    if (r != C) {
      const double mlt = a[r][C];
      if (mlt != 0.0) {
#pragma unroll
        for (int q = 0; q < 10; q++) a[r][q] -= mlt * a[C][q];
      }
    }
}
__device__ __forceinline__ bool inv5p1_flex(const double* f5, double f55, double* k5, double& k55) {
  double a[5][10];
#pragma unroll
  for (int r = 0; r < 5; r++)
#pragma unroll
    for (int c = 0; c < 5; c++) { a[r][c] = f5[r + 5 * c]; a[r][5 + c] = (r == c) ? 1.0 : 0.0; }
  bool ok = true;
  gj5_step<0>(a, ok); gj5_step<1>(a, ok); gj5_step<2>(a, ok); gj5_step<3>(a, ok); gj5_step<4>(a, ok);
#pragma unroll
  for (int r = 0; r < 5; r++)
#pragma unroll
    for (int c = 0; c < 5; c++) k5[r + 5 * c] = a[r][5 + c];
  k55 = 1.0 / f55;
  return ok;
}

template <int G>
__global__ void __launch_bounds__(128, XB_FBC_SEC_OCC) fbc3d_update_sec_kernel(BeamView B, const double* __restrict__ U,
                                                               const double* __restrict__ DU, int* fail) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long le = tid / G;
  const int i = (int)(tid - le * G);         // this lane's section
  if (le >= (B.ulist != nullptr ? B.nlist : B.n)) return;      // a whole group leaves together
  const long long e = B.ulist != nullptr ? (long long)B.ulist[le] : le;      // listed update: see GroupView::ulist
  const int lane = threadIdx.x & 31;
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
  const long long n = B.n;
  const int nip = B.nip;
  const bool act = i < nip;
  const int is = act ? i : 0;                // idle lanes shadow section 0's addresses, never store
  double R[9];
  const double L = B.geo[e];
#pragma unroll
  for (int q = 0; q < 9; q++) R[q] = B.geo[(size_t)(1 + q) * n + e];
  double v[6], dv[6], vin[6];
  {
    double ug[12], dug[12];
#pragma unroll
    for (int a = 0; a < 2; a++) {
      const int nd = B.conn[e * 2 + a];
#pragma unroll
      for (int j = 0; j < 6; j++) { ug[a * 6 + j] = U[(size_t)nd * 6 + j]; dug[a * 6 + j] = DU[(size_t)nd * 6 + j]; }
    }
    fbc3d_end_disp(B, e, ug); fbc3d_end_disp(B, e, dug);
    crd3d_basic(L, R, ug, v);
    crd3d_basic(L, R, dug, dv);
    if (B.pdelta && act && i == 0) {   // crdTransf->update(): PDeltaCrdTransf3d.cpp:200-249 (before any early return)
      const double ul1 = R[3] * ug[0] + R[4] * ug[1] + R[5] * ug[2], ul2 = R[6] * ug[0] + R[7] * ug[1] + R[8] * ug[2];
      const double ul7 = R[3] * ug[6] + R[4] * ug[7] + R[5] * ug[8], ul8 = R[6] * ug[6] + R[7] * ug[7] + R[8] * ug[8];
      B.ul[e] = ul1 - ul7; B.ul[n + e] = ul2 - ul8;
    }
  }
  const int initialFlag = B.iflag[e];
  // (all lanes of an element take the same way out: dv and the load flag are the element's)
  // numEleLoads > 0 for THIS element (ForceBeamColumn3d.cpp: the early return needs numEleLoads == 0)
  const bool pointed = B.wl != nullptr && B.loads_on && B.has_point && (B.wl[3 * n + e] != 0.0 || B.wl[4 * n + e] != 0.0 || B.wl[5 * n + e] != 0.0);
  const bool partial = B.wl != nullptr && B.loads_on && B.has_partial && B.wl[12 * n + e] > B.wl[11 * n + e];
  const bool loaded = pointed || partial || (B.wl != nullptr && B.loads_on && (B.wl[e] != 0.0 || B.wl[n + e] != 0.0 || B.wl[2 * n + e] != 0.0));
  if (initialFlag != 0 && norm6(dv) <= DBL_EPSILON && !loaded) return;
#pragma unroll
  for (int q = 0; q < 6; q++) vin[q] = v[q] - dv[q];
  const double xL = rule_x(B, e, is), xL1 = xL - 1.0, wtL = rule_w(B, e, is) * L;
  // `eleLoad -beamUniform`: this section's forces sp (computeSectionForces, ForceBeamColumn3d.cpp:1197-1215)
  double sp0 = 0.0, sp1 = 0.0, sp2 = 0.0;
  if (loaded) {
    const double x = xL * L;
    const double wy = B.wl[e] * B.lam, wz = B.wl[n + e] * B.lam, wa = B.wl[2 * n + e] * B.lam;
    sp0 = wa * (L - x); sp1 = wy * 0.5 * x * (x - L); sp2 = wz * 0.5 * x * (L - x);
    if (pointed) {   // Beam3dPointLoad (ForceBeamColumn3d.cpp:1314-1373)
      const double Py = B.wl[3 * n + e] * B.lam, Pz = B.wl[4 * n + e] * B.lam, N = B.wl[5 * n + e] * B.lam, aOverL = B.wl[6 * n + e];
      const double a = aOverL * L;
      const double Vy1 = Py * (1.0 - aOverL), Vy2 = Py * aOverL, Vz1 = Pz * (1.0 - aOverL), Vz2 = Pz * aOverL;
      if (x <= a) { sp0 += N; sp1 -= x * Vy1; sp2 += x * Vz1; }
      else { sp1 -= (L - x) * Vy2; sp2 += (L - x) * Vz2; }
    }
    if (partial) {   // Beam3dPartialUniformLoad (ForceBeamColumn3d.cpp:1224-1313)
      const double wya = B.wl[7 * n + e] * B.lam, wyb = B.wl[8 * n + e] * B.lam, waa = B.wl[9 * n + e] * B.lam, wab = B.wl[10 * n + e] * B.lam;
      const double wza = B.wl[13 * n + e] * B.lam, wzb = B.wl[14 * n + e] * B.lam;
      const double a = B.wl[11 * n + e] * L, b = B.wl[12 * n + e] * L;
      const double Fa = waa * (b - a) + 0.5 * (wab - waa) * (b - a);
      double Fy = wya * (b - a), Fz = wza * (b - a);
      double c = a + 0.5 * (b - a);
      double VyI = Fy * (1 - c / L), VyJ = Fy * c / L, VzI = Fz * (1 - c / L), VzJ = Fz * c / L;
      Fy = 0.5 * (wyb - wya) * (b - a); Fz = 0.5 * (wzb - wza) * (b - a);
      c = a + 2.0 / 3.0 * (b - a);
      VyI += Fy * (1 - c / L); VyJ += Fy * c / L; VzI += Fz * (1 - c / L); VzJ += Fz * c / L;
      if (x <= a) { sp0 += Fa; sp1 -= VyI * x; sp2 += VzI * x; }
      else if (x >= b) { sp1 += VyJ * (x - L); sp2 -= VzJ * (x - L); }
      else {
        const double wyy = wya + (wyb - wya) / (b - a) * (x - a), wzz = wza + (wzb - wza) / (b - a) * (x - a);
        sp0 += Fa - waa * (x - a) - 0.5 * (wab - waa) / (b - a) * (x - a) * (x - a);
        sp1 += -VyI * x + 0.5 * wya * (x - a) * (x - a) + 0.5 * (wyy - wya) * (x - a) * (x - a) / 3.0;
        sp2 += VzI * x - 0.5 * wza * (x - a) * (x - a) - 0.5 * (wzz - wza) * (x - a) * (x - a) / 3.0;
      }
    }
  }
  // initial section flexibility: 3x3 block (column-major, stride 3) + torsion
  double f0[9], f0t;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) f0[r + 3 * c] = __ldg(B.fs0 + r + 4 * c);
  f0t = __ldg(B.fs0 + 15);
  double dvTrial[6], dvToDo[6];
#pragma unroll
  for (int q = 0; q < 6; q++) { dvToDo[q] = dv[q]; dvTrial[q] = dvToDo[q]; }
  int numSubdivide = 1;
  bool converged = false;
  const double factor = 10.0;
  const int maxSubdivisions = 10;
  int passes = 0;
  while (!converged && numSubdivide <= maxSubdivisions) {
    for (int l = 0; l < 3; l++) {
      double SeTrial[6], k5[25], k55;
#pragma unroll
      for (int q = 0; q < 6; q++) SeTrial[q] = B.Se[(size_t)q * n + e];
#pragma unroll
      for (int c = 0; c < 5; c++)
#pragma unroll
        for (int r = 0; r < 5; r++) k5[r + 5 * c] = B.kv[(size_t)(r + 6 * c) * n + e];
      k55 = B.kv[(size_t)35 * n + e];
      double vsS[4], SsrS[4], fsS[9], fsT;
#pragma unroll
      for (int q = 0; q < 4; q++) { vsS[q] = B.vs[(size_t)(is * 4 + q) * n + e]; SsrS[q] = B.Ssr[(size_t)(is * 4 + q) * n + e]; }
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) fsS[r + 3 * c] = B.fs[(size_t)(is * 16 + r + 4 * c) * n + e];
      fsT = B.fs[(size_t)(is * 16 + 15) * n + e];
      {
        // SeTrial += kv * dvTrial: column by column, as Vector::addMatrixVector does; the torsion row and
        // column of kv hold exact zeros off the diagonal
        double dSe[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 5; c++)
#pragma unroll
          for (int r = 0; r < 5; r++) dSe[r] += k5[r + 5 * c] * dvTrial[c];
        dSe[5] += k55 * dvTrial[5];
#pragma unroll
        for (int q = 0; q < 6; q++) SeTrial[q] += dSe[q];
      }
      int numIters = B.maxIters;
      if (l == 1) numIters = 10 * B.maxIters;
      for (int j = 0; j < numIters; j++) {
        if (++passes > XB_FBC3D_MAX_PASSES) { if (i == 0) atomicExch(fail, 2); return; }
        double fl5[25], fl55 = 0.0, vrl[6];
#pragma unroll
        for (int q = 0; q < 25; q++) fl5[q] = 0.0;
#pragma unroll
        for (int q = 0; q < 6; q++) vrl[q] = 0.0;
        if (act) {
          double Ss[4], dSs[4], dvs[4], ssec[4], ksec[16];
          Ss[0] = SeTrial[0];
          Ss[1] = xL1 * SeTrial[1] + xL * SeTrial[2];
          Ss[2] = xL1 * SeTrial[3] + xL * SeTrial[4];
          Ss[3] = SeTrial[5];
          if (loaded) { Ss[0] += sp0; Ss[1] += sp1; Ss[2] += sp2; }
#pragma unroll
          for (int q = 0; q < 4; q++) dSs[q] = Ss[q] - SsrS[q];
          const bool initial = (l == 1) || (l == 2 && j == 0);
          // dvs = fs * dSs (Vector::addMatrixVector, column by column; the zero couplings of the torsion
          // row / column add exact zeros in the reference)
#pragma unroll
          for (int q = 0; q < 4; q++) dvs[q] = 0.0;
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) dvs[r] += (initial ? f0[r + 3 * c] : fsS[r + 3 * c]) * dSs[c];
          dvs[3] += (initial ? f0t : fsT) * dSs[3];
          if (initialFlag != 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) vsS[q] += dvs[q];
          }
          section3_trial(B, e, i, vsS, ssec, ksec);
#pragma unroll
          for (int q = 0; q < 4; q++) SsrS[q] = ssec[q];
          {
            double a3[9];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
              for (int r = 0; r < 3; r++) a3[r + 3 * c] = ksec[r + 4 * c];
            inv3(a3, fsS);
            fsT = 1.0 / ksec[15];
          }
#pragma unroll
          for (int q = 0; q < 4; q++) dSs[q] = Ss[q] - SsrS[q];
#pragma unroll
          for (int q = 0; q < 4; q++) dvs[q] = 0.0;
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) dvs[r] += fsS[r + 3 * c] * dSs[c];
          dvs[3] += fsT * dSs[3];
          // fb = fs * b * wtL (4 x 6, only the P-Mz-My rows x the 5 flexural / axial columns are non-zero,
          // plus the torsion entry); this section's addend to f = b^T fb
          double fb[3][5];
#pragma unroll
          for (int r = 0; r < 3; r++) {
            fb[r][0] = fsS[r + 3 * 0] * wtL;
            { const double tmp = fsS[r + 3 * 1] * wtL; fb[r][1] = xL1 * tmp; fb[r][2] = xL * tmp; }
            { const double tmp = fsS[r + 3 * 2] * wtL; fb[r][3] = xL1 * tmp; fb[r][4] = xL * tmp; }
          }
#pragma unroll
          for (int c = 0; c < 5; c++) {
            fl5[0 + 5 * c] = fb[0][c];
            { const double tmp = fb[1][c]; fl5[1 + 5 * c] = xL1 * tmp; fl5[2 + 5 * c] = xL * tmp; }
            { const double tmp = fb[2][c]; fl5[3 + 5 * c] = xL1 * tmp; fl5[4 + 5 * c] = xL * tmp; }
          }
          fl55 = fsT * wtL;
#pragma unroll
          for (int q = 0; q < 4; q++) dvs[q] += vsS[q];
          { const double dei = dvs[0] * wtL; vrl[0] = dei; }
          { const double dei = dvs[1] * wtL; vrl[1] = xL1 * dei; vrl[2] = xL * dei; }
          { const double dei = dvs[2] * wtL; vrl[3] = xL1 * dei; vrl[4] = xL * dei; }
          { const double dei = dvs[3] * wtL; vrl[5] = dei; }
        }
        double f5[25], f55, vr[6];
#pragma unroll
        for (int q = 0; q < 25; q++) f5[q] = group_sum_ordered(fl5[q], gmask, gbase, nip);
        f55 = group_sum_ordered(fl55, gmask, gbase, nip);
#pragma unroll
        for (int q = 0; q < 6; q++) vr[q] = group_sum_ordered(vrl[q], gmask, gbase, nip);
        if (!inv5p1_flex(f5, f55, k5, k55)) { if (i == 0) atomicExch(fail, 2); return; }
#pragma unroll
        for (int q = 0; q < 6; q++) { dv[q] = vin[q]; dv[q] += dvTrial[q]; dv[q] -= vr[q]; }
        double dSe[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 5; c++)
#pragma unroll
          for (int r = 0; r < 5; r++) dSe[r] += k5[r + 5 * c] * dv[c];
        dSe[5] += k55 * dv[5];
        double dW = 0.0;
#pragma unroll
        for (int q = 0; q < 6; q++) dW += dv[q] * dSe[q];
#pragma unroll
        for (int q = 0; q < 6; q++) SeTrial[q] += dSe[q];
        if (fabs(dW) < B.tol) {
#pragma unroll
          for (int q = 0; q < 6; q++) { dvToDo[q] -= dvTrial[q]; vin[q] += dvTrial[q]; }
          if (norm6(dvToDo) <= DBL_EPSILON) converged = true;
          else {
#pragma unroll
            for (int q = 0; q < 6; q++) dvTrial[q] = dvToDo[q];
            numSubdivide = 1;
          }
          if (i == 0) {
#pragma unroll
            for (int q = 0; q < 6; q++) B.Se[(size_t)q * n + e] = SeTrial[q];
#pragma unroll
            for (int c = 0; c < 5; c++)
#pragma unroll
              for (int r = 0; r < 5; r++) B.kv[(size_t)(r + 6 * c) * n + e] = k5[r + 5 * c];
            B.kv[(size_t)35 * n + e] = k55;
          }
          if (act) {
#pragma unroll
            for (int q = 0; q < 4; q++) { B.vs[(size_t)(i * 4 + q) * n + e] = vsS[q]; B.Ssr[(size_t)(i * 4 + q) * n + e] = SsrS[q]; }
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
              for (int r = 0; r < 3; r++) B.fs[(size_t)(i * 16 + r + 4 * c) * n + e] = fsS[r + 3 * c];
            B.fs[(size_t)(i * 16 + 15) * n + e] = fsT;
          }
          // the next l iteration (if any) re-reads the state just stored: make it visible to the group
          __syncwarp(gmask);
          j = numIters + 1; l = 4;
        } else {
          if (j == (numIters - 1) && (l == 2)) {
#pragma unroll
            for (int q = 0; q < 6; q++) dvTrial[q] /= factor;
            numSubdivide++;
          }
        }
      }
    }
  }
  if (!converged) { if (i == 0) atomicExch(fail, 2); return; }
  if (i == 0) B.iflag[e] = 1;
}

// ForceBeamColumn2d::update with one lane per section (see fbc3d_update_sec_kernel)
template <int G>
__global__ void __launch_bounds__(128, XB_FBC_SEC_OCC) fbc2d_update_sec_kernel(BeamView B, const double* __restrict__ U,
                                                                               const double* __restrict__ DU, int* fail) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long le = tid / G;
  const int i = (int)(tid - le * G);
  if (le >= (B.ulist != nullptr ? B.nlist : B.n)) return;
  const long long e = B.ulist != nullptr ? (long long)B.ulist[le] : le;
  const int lane = threadIdx.x & 31;
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
  const long long n = B.n;
  const int nip = B.nip;
  const bool act = i < nip;
  const int is = act ? i : 0;
  const double L = B.geo[e], cosT = B.geo[n + e], sinT = B.geo[2 * n + e];
  double v[3], dv[3], vin[3];
  {
    double ug[6], dug[6];
#pragma unroll
    for (int a = 0; a < 2; a++) {
      const int nd = B.conn[e * 2 + a];
#pragma unroll
      for (int j = 0; j < 3; j++) { ug[a * 3 + j] = U[(size_t)nd * 3 + j]; dug[a * 3 + j] = DU[(size_t)nd * 3 + j]; }
    }
    fbc2d_end_disp(B, e, ug); fbc2d_end_disp(B, e, dug);
    if (B.corot) {   // crdTransf->update(); getBasicIncrDeltaDisp = ub - ubpr (CorotCrdTransf2d.cpp:364-371), before any early return
      double cg[3];
      if (!corot2d_geom(L, cosT, sinT, ug, cg, v)) { if (i == 0) atomicExch(fail, 2); return; }
#pragma unroll
      for (int q = 0; q < 3; q++) dv[q] = v[q] - B.ul[(size_t)q * n + e];
      __syncwarp(gmask);      // every lane of the element has read ubpr
      if (i == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++) B.ul[(size_t)q * n + e] = v[q];
      }
    } else {
      crd2d_basic(L, cosT, sinT, ug, v);
      crd2d_basic(L, cosT, sinT, dug, dv);
    }
  }
  const int initialFlag = B.iflag[e];
  // numEleLoads > 0 for THIS element (ForceBeamColumn2d.cpp:575)
  const bool pointed = B.wl != nullptr && B.loads_on && B.has_point && (B.wl[3 * n + e] != 0.0 || B.wl[5 * n + e] != 0.0);
  const bool partial = B.wl != nullptr && B.loads_on && B.has_partial && B.wl[12 * n + e] > B.wl[11 * n + e];
  const bool loaded = pointed || partial || (B.wl != nullptr && B.loads_on && (B.wl[e] != 0.0 || B.wl[2 * n + e] != 0.0));
  if (initialFlag != 0 && sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]) <= DBL_EPSILON && !loaded) return;
#pragma unroll
  for (int q = 0; q < 3; q++) vin[q] = v[q] - dv[q];
  const double xL = rule_x(B, e, is), xL1 = xL - 1.0, wtL = rule_w(B, e, is) * L;
  // `eleLoad -beamUniform`: this section's forces sp (computeSectionForces, ForceBeamColumn2d.cpp:1034-1070)
  double sp0 = 0.0, sp1 = 0.0;
  if (loaded) {
    const double x = xL * L;
    const double wy = B.wl[e] * B.lam, wa = B.wl[2 * n + e] * B.lam;
    sp0 = wa * (L - x); sp1 = wy * 0.5 * x * (x - L);
    if (pointed) {   // Beam2dPointLoad (ForceBeamColumn2d.cpp:1138-1181)
      const double P = B.wl[3 * n + e] * B.lam, N = B.wl[5 * n + e] * B.lam, aOverL = B.wl[6 * n + e];
      const double a = aOverL * L;
      const double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
      if (x <= a) { sp0 += N; sp1 -= x * V1; }
      else sp1 -= (L - x) * V2;
    }
    if (partial) {   // Beam2dPartialUniformLoad (ForceBeamColumn2d.cpp:1073-1137)
      const double wya = B.wl[7 * n + e] * B.lam, wyb = B.wl[8 * n + e] * B.lam, waa = B.wl[9 * n + e] * B.lam, wab = B.wl[10 * n + e] * B.lam;
      const double a = B.wl[11 * n + e] * L, b = B.wl[12 * n + e] * L;
      const double Fa = waa * (b - a) + 0.5 * (wab - waa) * (b - a);
      double Fy = wya * (b - a);
      double c = a + 0.5 * (b - a);
      double VI = Fy * (1 - c / L), VJ = Fy * c / L;
      Fy = 0.5 * (wyb - wya) * (b - a);
      c = a + 2.0 / 3.0 * (b - a);
      VI += Fy * (1 - c / L); VJ += Fy * c / L;
      if (x <= a) { sp0 += Fa; sp1 -= VI * x; }
      else if (x >= b) sp1 += VJ * (x - L);
      else {
        const double wx = wya + (wyb - wya) / (b - a) * (x - a);
        sp0 += Fa - waa * (x - a) - 0.5 * (wab - waa) / (b - a) * (x - a) * (x - a);
        sp1 += -VI * x + wya * (x - a) * 0.5 * (x - a) + 0.5 * (wx - wya) * (x - a) * (x - a) / 3.0;
      }
    }
  }
  double f0[4];
#pragma unroll
  for (int q = 0; q < 4; q++) f0[q] = __ldg(B.fs0 + q);
  double dvTrial[3], dvToDo[3];
#pragma unroll
  for (int q = 0; q < 3; q++) { dvToDo[q] = dv[q]; dvTrial[q] = dvToDo[q]; }
  int numSubdivide = 1;
  bool converged = false;
  const double factor = 10;
  const int maxSubdivisions = 4;
  while (!converged && numSubdivide <= maxSubdivisions) {
    for (int l = 0; l < 3; l++) {
      double SeTrial[3], kvTrial[9];
#pragma unroll
      for (int q = 0; q < 3; q++) SeTrial[q] = B.Se[(size_t)q * n + e];
#pragma unroll
      for (int q = 0; q < 9; q++) kvTrial[q] = B.kv[(size_t)q * n + e];
      double vsS[2], SsrS[2], fsS[4];
#pragma unroll
      for (int q = 0; q < 2; q++) { vsS[q] = B.vs[(size_t)(is * 2 + q) * n + e]; SsrS[q] = B.Ssr[(size_t)(is * 2 + q) * n + e]; }
#pragma unroll
      for (int q = 0; q < 4; q++) fsS[q] = B.fs[(size_t)(is * 4 + q) * n + e];
      {
        double dSe[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int r = 0; r < 3; r++) dSe[r] += kvTrial[r + 3 * c] * dvTrial[c];
#pragma unroll
        for (int q = 0; q < 3; q++) SeTrial[q] += dSe[q];
      }
      int numIters = B.maxIters;
      if (l == 1) numIters = 10 * B.maxIters;
      for (int j = 0; j < numIters; j++) {
        double fl[9], vrl[3];
#pragma unroll
        for (int q = 0; q < 9; q++) fl[q] = 0.0;
        vrl[0] = vrl[1] = vrl[2] = 0.0;
        if (act) {
          double Ss[2], dSs[2], dvs[2], fb[6], ssec[2], ksec[4];
          Ss[0] = SeTrial[0];
          Ss[1] = xL1 * SeTrial[1] + xL * SeTrial[2];
          if (loaded) { Ss[0] += sp0; Ss[1] += sp1; }
          dSs[0] = Ss[0] - SsrS[0]; dSs[1] = Ss[1] - SsrS[1];
          const bool initial = (l == 1) || (l == 2 && j == 0);
          dvs[0] = 0.0; dvs[1] = 0.0;
#pragma unroll
          for (int c = 0; c < 2; c++)
#pragma unroll
            for (int r = 0; r < 2; r++) dvs[r] += (initial ? f0[r + 2 * c] : fsS[r + 2 * c]) * dSs[c];
          if (initialFlag != 0) { vsS[0] += dvs[0]; vsS[1] += dvs[1]; }
          section_trial(B, e, i, vsS, ssec, ksec);
          SsrS[0] = ssec[0]; SsrS[1] = ssec[1];
          section_flex(B, ksec, fsS);
          dSs[0] = Ss[0] - SsrS[0]; dSs[1] = Ss[1] - SsrS[1];
          dvs[0] = 0.0; dvs[1] = 0.0;
#pragma unroll
          for (int c = 0; c < 2; c++)
#pragma unroll
            for (int r = 0; r < 2; r++) dvs[r] += fsS[r + 2 * c] * dSs[c];
          // fb = fs * b * wtL (2 x 3), this section's addend to f = b^T fb
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            fb[jj + 2 * 0] = fsS[jj + 2 * 0] * wtL;
            const double tmp = fsS[jj + 2 * 1] * wtL; fb[jj + 2 * 1] = xL1 * tmp; fb[jj + 2 * 2] = xL * tmp;
          }
#pragma unroll
          for (int jj = 0; jj < 3; jj++) {
            fl[0 + 3 * jj] = fb[0 + 2 * jj];
            const double tmp = fb[1 + 2 * jj]; fl[1 + 3 * jj] = xL1 * tmp; fl[2 + 3 * jj] = xL * tmp;
          }
          dvs[0] += vsS[0]; dvs[1] += vsS[1];
          { const double dei = dvs[0] * wtL; vrl[0] = dei; }
          { const double dei = dvs[1] * wtL; vrl[1] = xL1 * dei; vrl[2] = xL * dei; }
        }
        double f[9], vr[3];
#pragma unroll
        for (int q = 0; q < 9; q++) f[q] = group_sum_ordered(fl[q], gmask, gbase, nip);
#pragma unroll
        for (int q = 0; q < 3; q++) vr[q] = group_sum_ordered(vrl[q], gmask, gbase, nip);
        inv3(f, kvTrial);
#pragma unroll
        for (int q = 0; q < 3; q++) { dv[q] = vin[q]; dv[q] += dvTrial[q]; dv[q] -= vr[q]; }
        double dSe[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int r = 0; r < 3; r++) dSe[r] += kvTrial[r + 3 * c] * dv[c];
        double dW = 0.0;
#pragma unroll
        for (int q = 0; q < 3; q++) dW += dv[q] * dSe[q];
#pragma unroll
        for (int q = 0; q < 3; q++) SeTrial[q] += dSe[q];
        if (fabs(dW) < B.tol) {
#pragma unroll
          for (int q = 0; q < 3; q++) { dvToDo[q] -= dvTrial[q]; vin[q] += dvTrial[q]; }
          if (sqrt(dvToDo[0] * dvToDo[0] + dvToDo[1] * dvToDo[1] + dvToDo[2] * dvToDo[2]) <= DBL_EPSILON) converged = true;
          else {
#pragma unroll
            for (int q = 0; q < 3; q++) dvTrial[q] = dvToDo[q];
            numSubdivide = 1;
          }
          if (i == 0) {
#pragma unroll
            for (int q = 0; q < 3; q++) B.Se[(size_t)q * n + e] = SeTrial[q];
#pragma unroll
            for (int q = 0; q < 9; q++) B.kv[(size_t)q * n + e] = kvTrial[q];
          }
          if (act) {
#pragma unroll
            for (int q = 0; q < 2; q++) { B.vs[(size_t)(i * 2 + q) * n + e] = vsS[q]; B.Ssr[(size_t)(i * 2 + q) * n + e] = SsrS[q]; }
#pragma unroll
            for (int q = 0; q < 4; q++) B.fs[(size_t)(i * 4 + q) * n + e] = fsS[q];
          }
          __syncwarp(gmask);
          j = numIters + 1; l = 3;
        } else {
          if (j == (numIters - 1) && (l == 2)) {
#pragma unroll
            for (int q = 0; q < 3; q++) dvTrial[q] /= factor;
            numSubdivide++;
          }
        }
      }
    }
  }
  if (!converged) { if (i == 0) atomicExch(fail, 2); return; }
  if (i == 0) B.iflag[e] = 1;
}

// getTangentStiff -> LinearCrdTransf3d::getGlobalStiffMatrix(kv); getResistingForce ->
// getGlobalResistingForce(Se).  Rows of node a (6 of them) go to that node's slot.
__global__ void __launch_bounds__(64) fbc3d_form_kernel(BeamView B, int want_k, int want_r, int transpose, BeamDyn dy) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n) return;
  const long long n = B.n;
  const double L = B.geo[e], oneOverL = 1.0 / L;
  double R[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = B.geo[(size_t)(1 + 3 * r + c) * n + e];
  if (want_k) {
    double kb[6][6], kl[12][12], tmp[12][12];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) kb[i][j] = B.kv[(size_t)(i + 6 * j) * n + e];
    if (dy.k_on) {
      for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
        double t = dy.at * kb[i][j];
        if (dy.a0 != 0.0) t += dy.a0 * B.kv0[(size_t)(i + 6 * j) * n + e];
        if (dy.ac != 0.0) t += dy.ac * B.kvK[(size_t)(i + 6 * j) * n + e];
        kb[i][j] = t;
      }
    }
    for (int i = 0; i < 6; i++) {
      tmp[i][0] = -kb[i][0];
      tmp[i][1] = oneOverL * (kb[i][1] + kb[i][2]);
      tmp[i][2] = -oneOverL * (kb[i][3] + kb[i][4]);
      tmp[i][3] = -kb[i][5];
      tmp[i][4] = kb[i][3];
      tmp[i][5] = kb[i][1];
      tmp[i][6] = kb[i][0];
      tmp[i][7] = -tmp[i][1];
      tmp[i][8] = -tmp[i][2];
      tmp[i][9] = kb[i][5];
      tmp[i][10] = kb[i][4];
      tmp[i][11] = kb[i][2];
    }
    for (int i = 0; i < 12; i++) {
      kl[0][i] = -tmp[0][i];
      kl[1][i] = oneOverL * (tmp[1][i] + tmp[2][i]);
      kl[2][i] = -oneOverL * (tmp[3][i] + tmp[4][i]);
      kl[3][i] = -tmp[5][i];
      kl[4][i] = tmp[3][i];
      kl[5][i] = tmp[1][i];
      kl[6][i] = tmp[0][i];
      kl[7][i] = -kl[1][i];
      kl[8][i] = -kl[2][i];
      kl[9][i] = tmp[5][i];
      kl[10][i] = tmp[4][i];
      kl[11][i] = tmp[2][i];
    }
    if (B.pdelta) {   // PDeltaCrdTransf3d::getGlobalStiffMatrix, PDeltaCrdTransf3d.cpp:873-881
      const double NoverL = (dy.k_on ? dy.at * B.Se[e] + (dy.ac != 0.0 ? dy.ac * B.nK[e] : 0.0) : B.Se[e]) * oneOverL;
      kl[1][1] += NoverL; kl[2][2] += NoverL; kl[7][7] += NoverL; kl[8][8] += NoverL;
      kl[1][7] -= NoverL; kl[7][1] -= NoverL; kl[2][8] -= NoverL; kl[8][2] -= NoverL;
    }
    for (int m = 0; m < 12; m++)
      for (int blk = 0; blk < 4; blk++)
        for (int c = 0; c < 3; c++)
          tmp[m][3 * blk + c] = kl[m][3 * blk] * R[0][c] + kl[m][3 * blk + 1] * R[1][c] + kl[m][3 * blk + 2] * R[2][c];
    // kg(3 blk + c, m) = sum_r R[r][c] tmp[3 blk + r][m]; kl is free now: reuse it for kg
    for (int m = 0; m < 12; m++)
      for (int blk = 0; blk < 4; blk++)
        for (int c = 0; c < 3; c++)
          kl[3 * blk + c][m] = R[0][c] * tmp[3 * blk][m] + R[1][c] * tmp[3 * blk + 1][m] + R[2][c] * tmp[3 * blk + 2][m];
    if (B.off) {   // K_node = To' K_end To (the joint-offset terms of LinearCrdTransf3d::getGlobalStiffMatrix)
      for (int a = 0; a < 2; a++) {
        const double dx = B.off[(size_t)(3 * a) * n + e], dy = B.off[(size_t)(3 * a + 1) * n + e], dz = B.off[(size_t)(3 * a + 2) * n + e];
        const double C[3][3] = {{0.0, dz, -dy}, {-dz, 0.0, dx}, {dy, -dx, 0.0}};      // u_end = u + C theta
        for (int i = 0; i < 12; i++) for (int j = 0; j < 3; j++)
          kl[i][6 * a + 3 + j] += kl[i][6 * a] * C[0][j] + kl[i][6 * a + 1] * C[1][j] + kl[i][6 * a + 2] * C[2][j];
        for (int i = 0; i < 12; i++) for (int j = 0; j < 3; j++)
          kl[6 * a + 3 + j][i] += C[0][j] * kl[6 * a][i] + C[1][j] * kl[6 * a + 1][i] + C[2][j] * kl[6 * a + 2][i];
      }
    }
    for (int a = 0; a < 2; a++) {
      const long long d = B.kdst[e * 2 + a];
      double* base = d >= 0 ? B.KeN + d : B.sendK + (-d - 1);
      for (int p = 0; p < 6; p++)
        for (int c = 0; c < 12; c++) base[p * B.cps + c] = transpose ? kl[c][a * 6 + p] : kl[a * 6 + p][c];
    }
  }
  if (want_r) {
    double q[6];
    for (int i = 0; i < 6; i++) q[i] = B.Se[i * n + e];
    const double q0s = q[0];
    double gdv1 = 0.0, gdv2 = 0.0;       // PDelta: geometric part of (bK Kt + bKc Kc) v on the local y and z dofs
    if (dy.r_on) {
      double vg[12], vb[6], Rf[9];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rf[3 * r + c] = R[r][c];
      for (int a = 0; a < 2; a++) {
        const int nd = B.conn[e * 2 + a];
        for (int j = 0; j < 6; j++) vg[a * 6 + j] = dy.V[(size_t)nd * 6 + j];
      }
      fbc3d_end_disp(B, e, vg);
      crd3d_basic(L, Rf, vg, vb);
      if (B.pdelta) {
        const double gd = (dy.bK * B.Se[e] + (dy.bKc != 0.0 ? dy.bKc * B.nK[e] : 0.0)) * oneOverL;
        gdv1 = gd * ((R[1][0] * vg[0] + R[1][1] * vg[1] + R[1][2] * vg[2]) - (R[1][0] * vg[6] + R[1][1] * vg[7] + R[1][2] * vg[8]));
        gdv2 = gd * ((R[2][0] * vg[0] + R[2][1] * vg[1] + R[2][2] * vg[2]) - (R[2][0] * vg[6] + R[2][1] * vg[7] + R[2][2] * vg[8]));
      }
      double qd[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) {
          double kd = dy.bK * B.kv[(size_t)(r + 6 * c) * n + e];
          if (dy.bK0 != 0.0) kd += dy.bK0 * B.kv0[(size_t)(r + 6 * c) * n + e];
          if (dy.bKc != 0.0) kd += dy.bKc * B.kvK[(size_t)(r + 6 * c) * n + e];
          qd[r] += kd * vb[c];
        }
      for (int i = 0; i < 6; i++) q[i] += qd[i];
    }
    double pl[12];
    pl[0] = -q[0]; pl[1] = oneOverL * (q[1] + q[2]); pl[2] = -oneOverL * (q[3] + q[4]); pl[3] = -q[5];
    pl[4] = q[3]; pl[5] = q[1]; pl[6] = q[0]; pl[7] = -pl[1]; pl[8] = -pl[2]; pl[9] = q[5]; pl[10] = q[4]; pl[11] = q[2];
    if (B.wl != nullptr && B.loads_on) {   // computeReactions (ForceBeamColumn3d.cpp:419-431) into LinearCrdTransf3d.cpp:727-731
      const double wy = B.wl[e] * B.lam, wz = B.wl[n + e] * B.lam, wa = B.wl[2 * n + e] * B.lam;
      double p0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      p0[0] -= wa * L;
      double Vr = 0.5 * wy * L;
      p0[1] -= Vr; p0[2] -= Vr;
      Vr = 0.5 * wz * L;
      p0[3] -= Vr; p0[4] -= Vr;
      if (B.has_point) {   // Beam3dPointLoad (ForceBeamColumn3d.cpp:457-475)
        const double Py = B.wl[3 * n + e] * B.lam, Pz = B.wl[4 * n + e] * B.lam, N = B.wl[5 * n + e] * B.lam, aOverL = B.wl[6 * n + e];
        double V1 = Py * (1.0 - aOverL), V2 = Py * aOverL;
        p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
        V1 = Pz * (1.0 - aOverL); V2 = Pz * aOverL;
        p0[3] -= V1; p0[4] -= V2;
      }
      if (B.has_partial && B.wl[12 * n + e] > B.wl[11 * n + e]) {   // Beam3dPartialUniformLoad (ForceBeamColumn3d.cpp:432-456)
        const double wya = B.wl[7 * n + e] * B.lam, wyb = B.wl[8 * n + e] * B.lam, waa = B.wl[9 * n + e] * B.lam, wab = B.wl[10 * n + e] * B.lam;
        const double wza = B.wl[13 * n + e] * B.lam, wzb = B.wl[14 * n + e] * B.lam;
        const double a = B.wl[11 * n + e] * L, b = B.wl[12 * n + e] * L;
        p0[0] -= waa * (b - a) + 0.5 * (wab - waa) * (b - a);
        double c = a + 0.5 * (b - a);
        double Fy = wya * (b - a);
        p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
        double Fz = wza * (b - a);
        p0[3] -= Fz * (1 - c / L); p0[4] -= Fz * c / L;
        c = a + 2.0 / 3.0 * (b - a);
        Fy = 0.5 * (wyb - wya) * (b - a);
        p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
        Fz = 0.5 * (wzb - wza) * (b - a);
        p0[3] -= Fz * (1 - c / L); p0[4] -= Fz * c / L;
      }
      pl[0] += p0[0]; pl[1] += p0[1]; pl[7] += p0[2]; pl[2] += p0[3]; pl[8] += p0[4];
    }
    if (B.pdelta) {   // PDeltaCrdTransf3d::getGlobalResistingForce, PDeltaCrdTransf3d.cpp:784-790
      double NoverL = B.ul[e] * q0s * oneOverL;
      pl[1] += NoverL; pl[7] -= NoverL;
      NoverL = B.ul[n + e] * q0s * oneOverL;
      pl[2] += NoverL; pl[8] -= NoverL;
      pl[1] += gdv1; pl[7] -= gdv1; pl[2] += gdv2; pl[8] -= gdv2;
    }
    double* Rg = B.Re + e * 12;
    for (int blk = 0; blk < 4; blk++)
      for (int c = 0; c < 3; c++)
        Rg[3 * blk + c] = R[0][c] * pl[3 * blk] + R[1][c] * pl[3 * blk + 1] + R[2][c] * pl[3 * blk + 2];
    if (B.off) {   // R_node = To' R_end: the end forces' moments about the nodes, d x F
      for (int a = 0; a < 2; a++) {
        const double dx = B.off[(size_t)(3 * a) * n + e], dy = B.off[(size_t)(3 * a + 1) * n + e], dz = B.off[(size_t)(3 * a + 2) * n + e];
        const double Fx = Rg[6 * a], Fy = Rg[6 * a + 1], Fz = Rg[6 * a + 2];
        Rg[6 * a + 3] += dy * Fz - dz * Fy; Rg[6 * a + 4] += dz * Fx - dx * Fz; Rg[6 * a + 5] += dx * Fy - dy * Fx;
      }
    }
  }
}

// ForceBeamColumn3d::revertToLastCommit (fibre records already copied back, committed -> trial)
__global__ void __launch_bounds__(64) fbc3d_revert_kernel(BeamView B) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n) return;
  const long long n = B.n;
  for (int i = 0; i < B.nip; i++) {
    double vs[4], s[4], k[16], fl[16];
    for (int q = 0; q < 4; q++) { vs[q] = B.vsc[(size_t)(i * 4 + q) * n + e]; B.vs[(size_t)(i * 4 + q) * n + e] = vs[q]; }
    section3_trial(B, e, i, vs, s, k);
    section3_flex(k, fl);
    for (int q = 0; q < 4; q++) B.Ssr[(size_t)(i * 4 + q) * n + e] = s[q];
    for (int q = 0; q < 16; q++) B.fs[(size_t)(i * 16 + q) * n + e] = fl[q];
  }
  for (int i = 0; i < 6; i++) B.Se[i * n + e] = B.Sec[i * n + e];
  for (int i = 0; i < 36; i++) B.kv[i * n + e] = B.kvc[i * n + e];
  B.iflag[e] = 0;
}

// ForceBeamColumn2d/3d::getInitialStiff in basic coordinates: the inverse of the initial flexibility
// sum_i b_i^T fs0 b_i w_i L (getInitialFlexibility); thread per element, run once when betaK0 is first set
__global__ void fbc_kv0_kernel(BeamView B) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n) return;
  const long long n = B.n;
  const double L = B.geo[e];
  if (B.nb == 3) {
    double f[9], k0[9], fS[4];
    for (int q = 0; q < 9; q++) f[q] = 0.0;
    for (int q = 0; q < 4; q++) fS[q] = __ldg(B.fs0 + q);
    for (int i = 0; i < B.nip; i++) {
      const double xL = rule_x(B, e, i), xL1 = xL - 1.0, wtL = rule_w(B, e, i) * L;
      double fb[6];
      for (int q = 0; q < 6; q++) fb[q] = 0.0;
      for (int jj = 0; jj < 2; jj++) fb[jj + 2 * 0] += fS[jj + 2 * 0] * wtL;
      for (int jj = 0; jj < 2; jj++) { const double tmp = fS[jj + 2 * 1] * wtL; fb[jj + 2 * 1] += xL1 * tmp; fb[jj + 2 * 2] += xL * tmp; }
      for (int jj = 0; jj < 3; jj++) f[0 + 3 * jj] += fb[0 + 2 * jj];
      for (int jj = 0; jj < 3; jj++) { const double tmp = fb[1 + 2 * jj]; f[1 + 3 * jj] += xL1 * tmp; f[2 + 3 * jj] += xL * tmp; }
    }
    inv3(f, k0);
    for (int q = 0; q < 9; q++) B.kv0[(size_t)q * n + e] = k0[q];
  } else {
    double f5[25], f55 = 0.0, k5[25], k55;
    for (int q = 0; q < 25; q++) f5[q] = 0.0;
    double fS[9];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) fS[r + 3 * c] = __ldg(B.fs0 + r + 4 * c);
    const double fT = __ldg(B.fs0 + 15);
    for (int i = 0; i < B.nip; i++) {
      const double xL = rule_x(B, e, i), xL1 = xL - 1.0, wtL = rule_w(B, e, i) * L;
      double fb[3][5];
      for (int r = 0; r < 3; r++) {
        fb[r][0] = fS[r + 3 * 0] * wtL;
        { const double tmp = fS[r + 3 * 1] * wtL; fb[r][1] = xL1 * tmp; fb[r][2] = xL * tmp; }
        { const double tmp = fS[r + 3 * 2] * wtL; fb[r][3] = xL1 * tmp; fb[r][4] = xL * tmp; }
      }
      for (int c = 0; c < 5; c++) {
        f5[0 + 5 * c] += fb[0][c];
        { const double tmp = fb[1][c]; f5[1 + 5 * c] += xL1 * tmp; f5[2 + 5 * c] += xL * tmp; }
        { const double tmp = fb[2][c]; f5[3 + 5 * c] += xL1 * tmp; f5[4 + 5 * c] += xL * tmp; }
      }
      f55 += fT * wtL;
    }
    inv5p1_flex(f5, f55, k5, k55);
    for (int q = 0; q < 36; q++) B.kv0[(size_t)q * n + e] = 0.0;
    for (int c = 0; c < 5; c++) for (int r = 0; r < 5; r++) B.kv0[(size_t)(r + 6 * c) * n + e] = k5[r + 5 * c];
    B.kv0[(size_t)35 * n + e] = k55;
  }
}

// Node::setTrialDisp bookkeeping: DU = Unew - U (incrDeltaDisp), U = Unew
__global__ void set_disp_kernel(long long ndof, const double* __restrict__ Unew, double* __restrict__ U,
                                double* __restrict__ DU) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndof) return;
  const double u = Unew[i];
  DU[i] = u - U[i];
  U[i] = u;
}

}  // namespace xbk
