// Device-side constitutive and shape-function routines shared by the element kernels.
// sm_100a only.  FP64 throughout (BASELINE.json: 1e-12 relative parity with the CPU path).
//
// Reference counterparts (under /root/reference/SRC):
//   shp3d                      interpolate/shp3d.cpp:33-168
//   FourNodeQuad::shapeFunction element/Plane/FourNodeQuad.cpp:1128-1196
//   J2Plasticity::plastic_integrator material/plastic/J2Plasticity.cpp:231-405
//   ElasticIsotropicThreeDimensional::getStress/getTangent material/elastic/...ThreeDimensional.cpp:88-128
//   ElasticIsotropicPlaneStrain2D                       material/elastic/...PlaneStrain2D.cpp:100-131
#pragma once
#include <cuda_runtime.h>

namespace xbk {

#define XB_HD __host__ __device__ __forceinline__

// symmetric 6x6 packed index (upper triangle, row-major): 21 entries
XB_HD constexpr int sym6(int i, int j) {
  return i <= j ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j));
}

// ---- rank-4 identity tensors in the reference's 6-vector order (00,11,22,01,12,20) ----
// matrix/identity.h: IbunI = I (x) I, IIdev = deviatoric projector
XB_HD constexpr double IbunI6(int a, int b) { return (a < 3 && b < 3) ? 1.0 : 0.0; }
XB_HD constexpr double IIdev6(int a, int b) {
  return (a < 3 && b < 3) ? (a == b ? 2. / 3. : -1. / 3.) : ((a == b) ? 0.5 : 0.0);
}

// J2 tangent entry from the stored compact form (normal n[6], c2, c3), same operation order
// as J2Plasticity.cpp:370-383
XB_HD double j2_tangent_entry(int a, int b, double bulk, double shear, const double* n, double c2,
                               double c3) {
  double NbunN = n[a] * n[b];
  double t = bulk * IbunI6(a, b);
  t += (2.0 * shear) * IIdev6(a, b);
  t += c2 * NbunN;
  t += c3 * (IIdev6(a, b) - NbunN);
  return t;
}

struct J2Result {
  double ep[6];   // epsilon_p_nplus1 (00,11,22,01,12,20)
  double xi;      // xi_nplus1
  double sig[6];  // stress tensor components (00,11,22,01,12,20)
  double nrm[6];  // normal
  double c2, c3;  // plastic tangent coefficients (c1*theta_inv, c1*gamma*inv_norm_tau)
  int fail;
};

// q(xi), q'(xi): J2Plasticity.cpp:437-448
XB_HD double j2_q(const double* p, double xi) { return p[5] * xi + p[3] + (p[2] - p[3]) * exp(-p[4] * xi); }
XB_HD double j2_qprime(const double* p, double xi) { return (p[2] - p[3]) * (-p[4]) * exp(-p[4] * xi) + p[5]; }

// par = K, G, sigma_0, sigma_infty, delta, H, eta ; e = strain TENSOR components
// (00,11,22,01,12,20) i.e. shear entries already halved (J2ThreeDimensional.cpp:118-136).
// dt = ops_Dt (0 in static analysis: the viscous terms drop out, J2Plasticity.cpp:283).
XB_HD void j2_integrate(const double* p, const double* e, const double* epn, double xin, double dt,
                        J2Result& r) {
  const double bulk = p[0], shear = p[1], sigma_0 = p[2], eta = p[6];
  const double root23 = 0.81649658092772603;  // sqrt(2/3)
  const double tolerance = 1.0e-10 * sigma_0;
  const double trace = e[0] + e[1] + e[2];
  double dev[6], ds[6];
  dev[0] = e[0] - 1. / 3. * trace; dev[1] = e[1] - 1. / 3. * trace; dev[2] = e[2] - 1. / 3. * trace;
  dev[3] = e[3]; dev[4] = e[4]; dev[5] = e[5];
#pragma unroll
  for (int c = 0; c < 6; c++) ds[c] = (dev[c] - epn[c]) * (2.0 * shear);
  // sum over the full 3x3 tensor in the reference's row-major order
  double nt = ds[0] * ds[0];
  nt += ds[3] * ds[3]; nt += ds[5] * ds[5];
  nt += ds[3] * ds[3]; nt += ds[1] * ds[1]; nt += ds[4] * ds[4];
  nt += ds[5] * ds[5]; nt += ds[4] * ds[4]; nt += ds[2] * ds[2];
  const double norm_tau = sqrt(nt);
  double inv_norm_tau = 0.0;
  if (norm_tau > tolerance) {
    inv_norm_tau = 1.0 / norm_tau;
#pragma unroll
    for (int c = 0; c < 6; c++) r.nrm[c] = inv_norm_tau * ds[c];
  } else {
#pragma unroll
    for (int c = 0; c < 6; c++) r.nrm[c] = 0.0;
  }
  const double phi = norm_tau - root23 * j2_q(p, xin);
  double gamma = 0.0, theta_inv = 0.0;
  r.fail = 0;
  if (phi > 0.0) {
    double resid = 1.0;
    int it = 0;
    const bool visc = (eta > 0.0 && dt > 0.0);
    while (fabs(resid) > tolerance) {
      resid = norm_tau - (2.0 * shear) * gamma - root23 * j2_q(p, xin + root23 * gamma);
      if (visc) resid -= (eta / dt) * gamma;
      double tang = -(2.0 * shear) - 2. / 3. * j2_qprime(p, xin + root23 * gamma);
      if (visc) tang -= (eta / dt);
      gamma -= (resid / tang);
      if (++it > 25) { r.fail = 1; break; }
    }
    gamma *= 1.0 - 1e-08;
#pragma unroll
    for (int c = 0; c < 6; c++) r.ep[c] = epn[c] + gamma * r.nrm[c];
    r.xi = xin + root23 * gamma;
#pragma unroll
    for (int c = 0; c < 6; c++) ds[c] = (2.0 * shear) * (dev[c] - r.ep[c]);
    double theta = (2.0 * shear) + 2. / 3. * j2_qprime(p, r.xi);
    if (visc) theta += (eta / dt);
    theta_inv = 1.0 / theta;
  } else {
#pragma unroll
    for (int c = 0; c < 6; c++) r.ep[c] = epn[c];
    r.xi = xin;
  }
#pragma unroll
  for (int c = 0; c < 6; c++) r.sig[c] = ds[c];
  r.sig[0] += bulk * trace; r.sig[1] += bulk * trace; r.sig[2] += bulk * trace;
  const double c1 = -4.0 * shear * shear;
  r.c2 = c1 * theta_inv;
  r.c3 = c1 * gamma * inv_norm_tau;
}

// ---- 8-node brick shape functions at Gauss point g (count = 4i+2j+k, Brick.cpp:757-780) ----
__device__ __forceinline__ void brick_shp(int g, const double (&xl)[3][8], double (&shp)[4][8],
                                          double& xsj) {
  const double a = 0.57735026918962573;  // 1/sqrt(3): Brick::one_over_root3, Brick.cpp:57
  const double x0 = (g & 4) ? a : -a, x1 = (g & 2) ? a : -a, x2 = (g & 1) ? a : -a;
  const double ap1 = 1.0 + x0, am1 = 1.0 - x0, ap2 = 1.0 + x1, am2 = 1.0 - x1, ap3 = 1.0 + x2,
               am3 = 1.0 - x2;
  { double c1 = 0.125 * am1 * am2, c2 = 0.125 * am2 * am3, c3 = 0.125 * am1 * am3;
    shp[0][0] = -c2; shp[0][1] = c2; shp[1][0] = -c3; shp[1][3] = c3;
    shp[2][0] = -c1; shp[2][4] = c1; shp[3][0] = c1 * am3; shp[3][4] = c1 * ap3; }
  { double c1 = 0.125 * ap1 * ap2, c2 = 0.125 * ap2 * ap3, c3 = 0.125 * ap1 * ap3;
    shp[0][7] = -c2; shp[0][6] = c2; shp[1][5] = -c3; shp[1][6] = c3;
    shp[2][2] = -c1; shp[2][6] = c1; shp[3][2] = c1 * am3; shp[3][6] = c1 * ap3; }
  { double c1 = 0.125 * am1 * ap2, c2 = 0.125 * am2 * ap3, c3 = 0.125 * am1 * ap3;
    shp[0][4] = -c2; shp[0][5] = c2; shp[1][4] = -c3; shp[1][7] = c3;
    shp[2][3] = -c1; shp[2][7] = c1; shp[3][3] = c1 * am3; shp[3][7] = c1 * ap3; }
  { double c1 = 0.125 * ap1 * am2, c2 = 0.125 * ap2 * am3, c3 = 0.125 * ap1 * am3;
    shp[0][3] = -c2; shp[0][2] = c2; shp[1][1] = -c3; shp[1][2] = c3;
    shp[2][1] = -c1; shp[2][5] = c1; shp[3][1] = c1 * am3; shp[3][5] = c1 * ap3; }
  double xs[3][3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    xs[j][0] = (xl[j][1] - xl[j][0]) * shp[0][1] + (xl[j][2] - xl[j][3]) * shp[0][2] +
               (xl[j][5] - xl[j][4]) * shp[0][5] + (xl[j][6] - xl[j][7]) * shp[0][6];
    xs[j][1] = (xl[j][2] - xl[j][1]) * shp[1][2] + (xl[j][3] - xl[j][0]) * shp[1][3] +
               (xl[j][6] - xl[j][5]) * shp[1][6] + (xl[j][7] - xl[j][4]) * shp[1][7];
    xs[j][2] = (xl[j][4] - xl[j][0]) * shp[2][4] + (xl[j][5] - xl[j][1]) * shp[2][5] +
               (xl[j][6] - xl[j][2]) * shp[2][6] + (xl[j][7] - xl[j][3]) * shp[2][7];
  }
  double ad[3][3];
  ad[0][0] = xs[1][1] * xs[2][2] - xs[1][2] * xs[2][1];
  ad[0][1] = xs[2][1] * xs[0][2] - xs[2][2] * xs[0][1];
  ad[0][2] = xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1];
  ad[1][0] = xs[1][2] * xs[2][0] - xs[1][0] * xs[2][2];
  ad[1][1] = xs[2][2] * xs[0][0] - xs[2][0] * xs[0][2];
  ad[1][2] = xs[0][2] * xs[1][0] - xs[0][0] * xs[1][2];
  ad[2][0] = xs[1][0] * xs[2][1] - xs[1][1] * xs[2][0];
  ad[2][1] = xs[2][0] * xs[0][1] - xs[2][1] * xs[0][0];
  ad[2][2] = xs[0][0] * xs[1][1] - xs[0][1] * xs[1][0];
  xsj = xs[0][0] * ad[0][0] + xs[0][1] * ad[1][0] + xs[0][2] * ad[2][0];
  const double rxsj = 1.0 / xsj;
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) xs[i][j] = ad[i][j] * rxsj;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double c1 = shp[0][k] * xs[0][0] + shp[1][k] * xs[1][0] + shp[2][k] * xs[2][0];
    double c2 = shp[0][k] * xs[0][1] + shp[1][k] * xs[1][1] + shp[2][k] * xs[2][1];
    double c3 = shp[0][k] * xs[0][2] + shp[1][k] * xs[1][2] + shp[2][k] * xs[2][2];
    shp[0][k] = c1; shp[1][k] = c2; shp[2][k] = c3;
  }
}

// ---- 4-node quad shape functions at (xi, eta); c = nodal coordinates [4][2]; returns detJ ----
__device__ __forceinline__ double quad_shp(double xi, double eta, const double (&c)[4][2],
                                           double (&shp)[3][4]) {
  const double oneMinuseta = 1.0 - eta, onePluseta = 1.0 + eta, oneMinusxi = 1.0 - xi,
               onePlusxi = 1.0 + xi;
  shp[2][0] = 0.25 * oneMinusxi * oneMinuseta;
  shp[2][1] = 0.25 * onePlusxi * oneMinuseta;
  shp[2][2] = 0.25 * onePlusxi * onePluseta;
  shp[2][3] = 0.25 * oneMinusxi * onePluseta;
  double J00 = 0.25 * (-c[0][0] * oneMinuseta + c[1][0] * oneMinuseta + c[2][0] * (onePluseta) -
                       c[3][0] * (onePluseta));
  double J01 = 0.25 * (-c[0][0] * oneMinusxi - c[1][0] * onePlusxi + c[2][0] * onePlusxi +
                       c[3][0] * oneMinusxi);
  double J10 = 0.25 * (-c[0][1] * oneMinuseta + c[1][1] * oneMinuseta + c[2][1] * onePluseta -
                       c[3][1] * onePluseta);
  double J11 = 0.25 * (-c[0][1] * oneMinusxi - c[1][1] * onePlusxi + c[2][1] * onePlusxi +
                       c[3][1] * oneMinusxi);
  const double detJ = J00 * J11 - J01 * J10;
  const double oneOverdetJ = 1.0 / detJ;
  const double L00 = 0.25 * (J11 * oneOverdetJ), L10 = 0.25 * (-J01 * oneOverdetJ),
               L01 = 0.25 * (-J10 * oneOverdetJ), L11 = 0.25 * (J00 * oneOverdetJ);
  const double L00oneMinuseta = L00 * oneMinuseta, L00onePluseta = L00 * onePluseta,
               L01oneMinusxi = L01 * oneMinusxi, L01onePlusxi = L01 * onePlusxi,
               L10oneMinuseta = L10 * oneMinuseta, L10onePluseta = L10 * onePluseta,
               L11oneMinusxi = L11 * oneMinusxi, L11onePlusxi = L11 * onePlusxi;
  shp[0][0] = -L00oneMinuseta - L01oneMinusxi;
  shp[0][1] = L00oneMinuseta - L01onePlusxi;
  shp[0][2] = L00onePluseta + L01onePlusxi;
  shp[0][3] = -L00onePluseta + L01oneMinusxi;
  shp[1][0] = -L10oneMinuseta - L11oneMinusxi;
  shp[1][1] = L10oneMinuseta - L11onePlusxi;
  shp[1][2] = L10onePluseta + L11onePlusxi;
  shp[1][3] = -L10onePluseta + L11oneMinusxi;
  return detJ;
}

// quadrature/Plane/LegendreFixedQuadrilateral.h:10-14 (15-digit literals, as in the reference)
__device__ __forceinline__ void quad_point(int i, double& xi, double& eta) {
  const double a = 0.577350269189626;
  xi = (i == 1 || i == 2) ? a : -a;
  eta = (i >= 2) ? a : -a;
}

}  // namespace xbk
