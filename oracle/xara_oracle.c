/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's
 * element-state-determination + assembly hot path.  Nothing under xara_b200/
 * may include, link or load this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks every function
 * below against the reference's own classes compiled from /root/reference
 * (oracle/ref_build.mk + oracle/ref_harness.cpp) and against the fixtures
 * under tests/golden/ that the same harness generated.
 *
 * Each function cites the reference file:line it restates.  The arithmetic
 * keeps the reference's operation order so that, built with
 * -ffp-contract=off, results agree with the reference to the last bits.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------ */
/* material kinds / element kinds shared with tests (mirrors include/xara_b200.h) */
enum { ORC_MAT_ELASTIC_ISOTROPIC = 0, ORC_MAT_J2 = 1 };
enum { ORC_ELE_BRICK = 0, ORC_ELE_QUAD = 1, ORC_ELE_FBC2D = 2, ORC_ELE_FBC3D = 3 };
enum { ORC_UNI_STEEL02 = 0, ORC_UNI_CONCRETE02 = 1, ORC_UNI_STEEL01 = 2, ORC_UNI_ELASTIC = 3, ORC_UNI_CONCRETE01 = 4, ORC_UNI_ELASTICPP = 5 };
enum { ORC_ND_3D = 0, ORC_ND_PLANE_STRAIN = 1, ORC_ND_PLANE_STRESS = 2 };

/* ======================================================================== */
/* J2Plasticity  (SRC/material/plastic/J2Plasticity.cpp)                     */
/* ======================================================================== */
typedef struct {
  /* parameters, J2Plasticity.h:123-130 */
  double bulk, shear, sigma_0, sigma_infty, delta, Hard, eta;
  /* internal variables, J2Plasticity.h:133-136 */
  double epsilon_p_n[3][3], epsilon_p_nplus1[3][3], xi_n, xi_nplus1;
  /* response, J2Plasticity.h:139-144 */
  double stress[3][3], tangent[3][3][3][3], strain[3][3];
  double commitEps22;   /* J2PlaneStress (material/Plane/J2PlaneStress.h): the out-of-plane strain at the last commit */
} OrcJ2;

/* SRC/matrix/identity.h: rank-4 I (x) I and the deviatoric projector */
static double IbunI(int i, int j, int k, int l) { return (i == j && k == l) ? 1.0 : 0.0; }
static double IIdev(int i, int j, int k, int l) {
  /* identity.h:42-77, entries are the literals 2./3., -1./3., 0.5 */
  if (i == j) {
    if (k == l) return (i == k) ? 2. / 3. : -1. / 3.;
    return 0.0;
  }
  if ((i == k && j == l) || (i == l && j == k)) return 0.5;
  return 0.0;
}

/* J2Plasticity.cpp:437-448 */
static double j2_q(const OrcJ2* m, double xi) {
  return m->Hard * xi + m->sigma_infty + (m->sigma_0 - m->sigma_infty) * exp(-m->delta * xi);
}
static double j2_qprime(const OrcJ2* m, double xi) {
  return (m->sigma_0 - m->sigma_infty) * (-m->delta) * exp(-m->delta * xi) + m->Hard;
}

/* J2Plasticity.cpp:452-504 (matrix index -> tensor indices) */
static void j2_index_map(int mi, int* i, int* j) {
  static const int I[6] = {0, 1, 2, 0, 1, 2}, J[6] = {0, 1, 2, 1, 2, 0};
  *i = I[mi]; *j = J[mi];
}

/* J2Plasticity.cpp:215-228 */
static void j2_zero(OrcJ2* m) {
  m->xi_n = m->xi_nplus1 = 0.0;
  memset(m->epsilon_p_n, 0, sizeof m->epsilon_p_n);
  memset(m->epsilon_p_nplus1, 0, sizeof m->epsilon_p_nplus1);
  memset(m->stress, 0, sizeof m->stress);
  memset(m->strain, 0, sizeof m->strain);
}

/* J2Plasticity::plastic_integrator, J2Plasticity.cpp:231-405.  dt = ops_Dt. */
static int j2_plastic_integrator(OrcJ2* m, double dt) {
  const double root23 = sqrt(2.0 / 3.0);
  const double tolerance = 1.0e-10 * m->sigma_0;
  const double shear = m->shear, bulk = m->bulk, eta = m->eta;
  double dev_stress[3][3], normal[3][3], dev_strain[3][3];
  double inv_norm_tau = 0.0, tang = 0.0;
  const int max_iterations = 25;
  int i, j;

  double trace = m->strain[0][0] + m->strain[1][1] + m->strain[2][2];
  memcpy(dev_strain, m->strain, sizeof dev_strain);
  for (i = 0; i < 3; i++) dev_strain[i][i] -= 1. / 3. * trace;

  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      dev_stress[i][j] = dev_strain[i][j];
      dev_stress[i][j] -= m->epsilon_p_n[i][j];
      dev_stress[i][j] *= 2.0 * shear;
    }

  double norm_tau = 0.0;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) norm_tau += dev_stress[i][j] * dev_stress[i][j];
  norm_tau = sqrt(norm_tau);

  if (norm_tau > tolerance) {
    inv_norm_tau = 1.0 / norm_tau;
    for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) normal[i][j] = inv_norm_tau * dev_stress[i][j];
  } else {
    memset(normal, 0, sizeof normal);
    inv_norm_tau = 0.0;
  }

  double phi = norm_tau - root23 * j2_q(m, m->xi_n);
  double c1, c2, c3, theta_inv = 0.0, gamma = 0.0;

  if (phi > 0.0) {
    double resid = 1.0;
    int iteration_counter = 0;
    while (fabs(resid) > tolerance) {
      resid = norm_tau - (2.0 * shear) * gamma - root23 * j2_q(m, m->xi_n + root23 * gamma);
      if (eta > 0.0 && dt > 0.0) resid -= (eta / dt) * gamma;
      tang = -(2.0 * shear) - 2. / 3. * j2_qprime(m, m->xi_n + root23 * gamma);
      if (eta > 0.0 && dt > 0.0) tang -= (eta / dt);
      gamma -= (resid / tang);
      iteration_counter++;
      if (iteration_counter > max_iterations) return -1;
    }
    gamma *= 1.0 - 1e-08;

    for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) m->epsilon_p_nplus1[i][j] = m->epsilon_p_n[i][j] + gamma * normal[i][j];
    m->xi_nplus1 = m->xi_n + root23 * gamma;

    for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) dev_stress[i][j] = (2.0 * shear) * (dev_strain[i][j] - m->epsilon_p_nplus1[i][j]);

    double theta = (2.0 * shear) + 2. / 3. * j2_qprime(m, m->xi_nplus1);
    if (eta > 0.0 && dt > 0.0) theta += (eta / dt);
    theta_inv = 1.0 / theta;
  } else {
    memcpy(m->epsilon_p_nplus1, m->epsilon_p_n, sizeof m->epsilon_p_n);
    m->xi_nplus1 = m->xi_n;
    gamma = 0.0;
    theta_inv = 0.0;
  }

  memcpy(m->stress, dev_stress, sizeof dev_stress);
  for (i = 0; i < 3; i++) m->stress[i][i] += bulk * trace;

  c1 = -4.0 * shear * shear;
  c2 = c1 * theta_inv;
  c3 = c1 * gamma * inv_norm_tau;

  for (int ii = 0; ii < 6; ii++)
    for (int jj = 0; jj < 6; jj++) {
      int k, l;
      j2_index_map(ii, &i, &j);
      j2_index_map(jj, &k, &l);
      double NbunN = normal[i][j] * normal[k][l];
      double t = bulk * IbunI(i, j, k, l);
      t += (2.0 * shear) * IIdev(i, j, k, l);
      t += c2 * NbunN;
      t += c3 * (IIdev(i, j, k, l) - NbunN);
      m->tangent[i][j][k][l] = t;
      m->tangent[j][i][k][l] = t;
      m->tangent[i][j][l][k] = t;
      m->tangent[j][i][l][k] = t;
    }
  return 0;
}

static void j2_init(OrcJ2* m, const double* p) {
  /* J2Plasticity.cpp:78-107 (full constructor) */
  m->bulk = p[0]; m->shear = p[1]; m->sigma_0 = p[2]; m->sigma_infty = p[3];
  m->delta = p[4]; m->Hard = p[5]; m->eta = p[6];
  j2_zero(m);
  j2_plastic_integrator(m, 0.0);
}

/* J2ThreeDimensional::setTrialStrain (J2ThreeDimensional.cpp:118-136) and
 * J2PlaneStrain::setTrialStrain (material/Plane/J2PlaneStrain.cpp:80-92) */
static int j2_set_trial_strain(OrcJ2* m, int type, const double* e, double dt) {
  if (type != ORC_ND_PLANE_STRESS) memset(m->strain, 0, sizeof m->strain);
  if (type == ORC_ND_3D) {
    m->strain[0][0] = e[0]; m->strain[1][1] = e[1]; m->strain[2][2] = e[2];
    m->strain[0][1] = 0.50 * e[3]; m->strain[1][0] = m->strain[0][1];
    m->strain[1][2] = 0.50 * e[4]; m->strain[2][1] = m->strain[1][2];
    m->strain[2][0] = 0.50 * e[5]; m->strain[0][2] = m->strain[2][0];
  } else if (type == ORC_ND_PLANE_STRESS) {
    /* J2PlaneStress::setTrialStrain (material/Plane/J2PlaneStress.cpp:123-180): the out-of-plane strain of the LAST
     * TRIAL is kept, then iterated until sigma_22 = 0 (|sigma_22| <= 1e-8 sigma_0, at most 26 integrator calls; the
     * integrator's return value is not looked at); the tangent is condensed in place */
    const double tolerance = 1.0e-8 * m->sigma_0;
    const int max_iterations = 25;
    int iteration_counter = 0;
    const double eps22 = m->strain[2][2];
    memset(m->strain, 0, sizeof m->strain);
    m->strain[0][0] = e[0]; m->strain[1][1] = e[1];
    m->strain[0][1] = 0.50 * e[2]; m->strain[1][0] = m->strain[0][1];
    m->strain[2][2] = eps22;
    do {
      j2_plastic_integrator(m, dt);
      m->strain[2][2] -= m->stress[2][2] / m->tangent[2][2][2][2];
      iteration_counter++;
      if (iteration_counter > max_iterations) break;
    } while (fabs(m->stress[2][2]) > tolerance);
    static const int PI[3] = {0, 1, 0}, PJ[3] = {0, 1, 1};
    for (int ii = 0; ii < 3; ii++)
      for (int jj = 0; jj < 3; jj++) {
        const int i = PI[ii], j = PJ[ii], k = PI[jj], l = PJ[jj];
        m->tangent[i][j][k][l] -= m->tangent[i][j][2][2] * m->tangent[2][2][k][l] / m->tangent[2][2][2][2];
        m->tangent[j][i][k][l] = m->tangent[i][j][k][l];
        m->tangent[i][j][l][k] = m->tangent[i][j][k][l];
        m->tangent[j][i][l][k] = m->tangent[i][j][k][l];
      }
    return 0;
  } else {
    m->strain[0][0] = e[0]; m->strain[1][1] = e[1];
    m->strain[0][1] = 0.50 * e[2]; m->strain[1][0] = m->strain[0][1];
  }
  return j2_plastic_integrator(m, dt);
}

/* getStress / getTangent: J2ThreeDimensional.cpp:183-226, J2PlaneStrain.cpp:118-146 */
static void j2_get_stress(const OrcJ2* m, int type, double* s) {
  if (type == ORC_ND_3D) {
    s[0] = m->stress[0][0]; s[1] = m->stress[1][1]; s[2] = m->stress[2][2];
    s[3] = m->stress[0][1]; s[4] = m->stress[1][2]; s[5] = m->stress[2][0];
  } else {
    s[0] = m->stress[0][0]; s[1] = m->stress[1][1]; s[2] = m->stress[0][1];
  }
}
static void j2_get_tangent(const OrcJ2* m, int type, double* D) {
  if (type == ORC_ND_3D) {
    for (int ii = 0; ii < 6; ii++)
      for (int jj = 0; jj < 6; jj++) {
        int i, j, k, l;
        j2_index_map(ii, &i, &j); j2_index_map(jj, &k, &l);
        D[ii * 6 + jj] = m->tangent[i][j][k][l];
      }
  } else {
    D[0] = m->tangent[0][0][0][0]; D[4] = m->tangent[1][1][1][1]; D[8] = m->tangent[0][1][0][1];
    D[1] = m->tangent[0][0][1][1]; D[3] = m->tangent[1][1][0][0];
    D[2] = m->tangent[0][0][0][1]; D[6] = m->tangent[0][1][0][0];
    D[5] = m->tangent[1][1][0][1]; D[7] = m->tangent[0][1][1][1];
  }
}
/* J2Plasticity::commitState, J2Plasticity.cpp:538-544 */
static void j2_commit(OrcJ2* m) {
  memcpy(m->epsilon_p_n, m->epsilon_p_nplus1, sizeof m->epsilon_p_n);
  m->xi_n = m->xi_nplus1;
}

/* ======================================================================== */
/* ElasticIsotropic (SRC/material/elastic/ElasticIsotropicThreeDimensional.cpp:88-128,
 *                   ElasticIsotropicPlaneStrain2D.cpp:100-131)                */
/* ======================================================================== */
typedef struct { double E, v; double epsilon[6], Cepsilon[6]; } OrcElastic;

static void el_get_tangent(const OrcElastic* m, int type, double* D) {
  double mu2 = m->E / (1.0 + m->v);
  double lam = m->v * mu2 / (1.0 - 2.0 * m->v);
  double mu = 0.50 * mu2;
  if (type == ORC_ND_PLANE_STRESS) {   /* ElasticIsotropicPlaneStress2D::getInitialTangent (material/elastic/ElasticIsotropicPlaneStress2D.cpp) */
    double d00 = m->E / (1.0 - m->v * m->v);
    double d01 = m->v * d00;
    double d22 = 0.5 * (d00 - d01);
    memset(D, 0, 9 * sizeof(double));
    D[0] = D[4] = d00; D[1] = D[3] = d01; D[8] = d22;
    return;
  }
  if (type == ORC_ND_3D) {
    memset(D, 0, 36 * sizeof(double));
    mu2 += lam;
    D[0] = D[7] = D[14] = mu2;
    D[1] = D[6] = D[2] = D[12] = D[8] = D[13] = lam;
    D[21] = mu; D[28] = mu; D[35] = mu;
  } else {
    memset(D, 0, 9 * sizeof(double));
    D[0] = D[4] = mu2 + lam;
    D[1] = D[3] = lam;
    D[8] = mu;
  }
}
static void el_get_stress(const OrcElastic* m, int type, double* s) {
  double mu2 = m->E / (1.0 + m->v);
  double lam = m->v * mu2 / (1.0 - 2.0 * m->v);
  double mu = 0.50 * mu2;
  mu2 += lam;
  const double* e = m->epsilon;
  if (type == ORC_ND_PLANE_STRESS) {   /* ElasticIsotropicPlaneStress2D::getStress */
    double d00 = m->E / (1.0 - m->v * m->v);
    double d01 = m->v * d00;
    double d22 = 0.5 * (d00 - d01);
    s[0] = d00 * e[0] + d01 * e[1];
    s[1] = d01 * e[0] + d00 * e[1];
    s[2] = d22 * e[2];
    return;
  }
  if (type == ORC_ND_3D) {
    s[0] = mu2 * e[0] + lam * (e[1] + e[2]);
    s[1] = mu2 * e[1] + lam * (e[0] + e[2]);
    s[2] = mu2 * e[2] + lam * (e[0] + e[1]);
    s[3] = mu * e[3]; s[4] = mu * e[4]; s[5] = mu * e[5];
  } else {
    s[0] = mu2 * e[0] + lam * e[1];
    s[1] = lam * e[0] + mu2 * e[1];
    s[2] = mu * e[2];
  }
}

/* ------------------------------------------------------------------------ */
/* one integration point of either kind */
typedef struct {
  int kind, type;
  double rho;        /* NDMaterial::getRho: J2Plasticity's rho (par[7]), ElasticIsotropicMaterial's rho (par[2]) */
  union { OrcJ2 j2; OrcElastic el; } u;
} OrcGP;

static void gp_init(OrcGP* g, int kind, int type, const double* p) {
  memset(g, 0, sizeof *g);
  g->kind = kind; g->type = type;
  if (kind == ORC_MAT_J2) { j2_init(&g->u.j2, p); g->rho = p[7]; }
  else { g->u.el.E = p[0]; g->u.el.v = p[1]; g->rho = p[2]; }
}
/* J2Plasticity::doInitialTangent (J2Plasticity.cpp:398-421) through J2ThreeDimensional::getInitialTangent
 * (J2ThreeDimensional.cpp:237) / J2PlaneStrain::getInitialTangent; elastic: the tangent itself */
static double j2_initial_entry(const OrcJ2* m, int i, int j, int k, int l) {
  double t = m->bulk * IbunI(i, j, k, l);
  t += (2.0 * m->shear) * IIdev(i, j, k, l);
  return t;
}
static void el_get_tangent(const OrcElastic* m, int type, double* D);
static void gp_get_initial_tangent(const OrcGP* g, double* D) {
  if (g->kind != ORC_MAT_J2) { el_get_tangent(&g->u.el, g->type, D); return; }
  const OrcJ2* m = &g->u.j2;
  if (g->type == ORC_ND_3D) {
    for (int ii = 0; ii < 6; ii++)
      for (int jj = 0; jj < 6; jj++) {
        int i, j, k, l;
        j2_index_map(ii, &i, &j); j2_index_map(jj, &k, &l);
        D[ii * 6 + jj] = j2_initial_entry(m, i, j, k, l);
      }
  } else {
    D[0] = j2_initial_entry(m, 0, 0, 0, 0); D[4] = j2_initial_entry(m, 1, 1, 1, 1); D[8] = j2_initial_entry(m, 0, 1, 0, 1);
    D[1] = j2_initial_entry(m, 0, 0, 1, 1); D[3] = j2_initial_entry(m, 1, 1, 0, 0);
    D[2] = j2_initial_entry(m, 0, 0, 0, 1); D[6] = j2_initial_entry(m, 0, 1, 0, 0);
    D[5] = j2_initial_entry(m, 1, 1, 0, 1); D[7] = j2_initial_entry(m, 0, 1, 1, 1);
  }
}
static int gp_set_trial_strain(OrcGP* g, const double* e) {
  int n = (g->type == ORC_ND_3D) ? 6 : 3;
  if (g->kind == ORC_MAT_J2) return j2_set_trial_strain(&g->u.j2, g->type, e, 0.0);
  memcpy(g->u.el.epsilon, e, n * sizeof(double));
  return 0;
}
static void gp_get_stress(const OrcGP* g, double* s) {
  if (g->kind == ORC_MAT_J2) j2_get_stress(&g->u.j2, g->type, s); else el_get_stress(&g->u.el, g->type, s);
}
static void gp_get_tangent(const OrcGP* g, double* D) {
  if (g->kind == ORC_MAT_J2) j2_get_tangent(&g->u.j2, g->type, D); else el_get_tangent(&g->u.el, g->type, D);
}
static void gp_commit(OrcGP* g) {
  if (g->kind == ORC_MAT_J2) { j2_commit(&g->u.j2); if (g->type == ORC_ND_PLANE_STRESS) g->u.j2.commitEps22 = g->u.j2.strain[2][2]; }   /* J2PlaneStress::commitState */
  else memcpy(g->u.el.Cepsilon, g->u.el.epsilon, sizeof g->u.el.epsilon);
}
static void gp_revert(OrcGP* g) {
  /* J2Plasticity::revertToLastCommit is a no-op (J2Plasticity.cpp:546-550);
   * ElasticIsotropicThreeDimensional.cpp:139-144 restores epsilon */
  if (g->kind != ORC_MAT_J2) memcpy(g->u.el.epsilon, g->u.el.Cepsilon, sizeof g->u.el.epsilon);
  else if (g->type == ORC_ND_PLANE_STRESS) g->u.j2.strain[2][2] = g->u.j2.commitEps22;   /* J2PlaneStress::revertToLastCommit */
}

/* material-level strain path; mirrors ref_nd_path in ref_harness.cpp */
int orc_nd_path(int kind, const double* p, int type, int n, const double* strains,
                const int* commit, double* stress, double* tangent) {
  OrcGP g; gp_init(&g, kind, type, p);
  int order = (type == ORC_ND_3D) ? 6 : 3;
  for (int s = 0; s < n; s++) {
    if (gp_set_trial_strain(&g, strains + (size_t)s * order) < 0) return -2;
    gp_get_stress(&g, stress + (size_t)s * order);
    gp_get_tangent(&g, tangent + (size_t)s * order * order);
    if (commit[s]) gp_commit(&g);
  }
  return order;
}

/* ======================================================================== */
/* shp3d  (SRC/interpolate/shp3d.cpp:33-168)                                  */
/* ======================================================================== */
static void shp3d(const double xn[3], double* xsj, double shp[4][8], const double xl[3][8]) {
  double ap1 = 1.0 + xn[0], am1 = 1.0 - xn[0];
  double ap2 = 1.0 + xn[1], am2 = 1.0 - xn[1];
  double ap3 = 1.0 + xn[2], am3 = 1.0 - xn[2];
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
  for (int j = 0; j < 3; j++) {
    xs[j][0] = (xl[j][1] - xl[j][0]) * shp[0][1] + (xl[j][2] - xl[j][3]) * shp[0][2]
             + (xl[j][5] - xl[j][4]) * shp[0][5] + (xl[j][6] - xl[j][7]) * shp[0][6];
    xs[j][1] = (xl[j][2] - xl[j][1]) * shp[1][2] + (xl[j][3] - xl[j][0]) * shp[1][3]
             + (xl[j][6] - xl[j][5]) * shp[1][6] + (xl[j][7] - xl[j][4]) * shp[1][7];
    xs[j][2] = (xl[j][4] - xl[j][0]) * shp[2][4] + (xl[j][5] - xl[j][1]) * shp[2][5]
             + (xl[j][6] - xl[j][2]) * shp[2][6] + (xl[j][7] - xl[j][3]) * shp[2][7];
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

  *xsj = xs[0][0] * ad[0][0] + xs[0][1] * ad[1][0] + xs[0][2] * ad[2][0];
  double rxsj = 1.0 / *xsj;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) xs[i][j] = ad[i][j] * rxsj;

  for (int k = 0; k < 8; k++) {
    double c1 = shp[0][k] * xs[0][0] + shp[1][k] * xs[1][0] + shp[2][k] * xs[2][0];
    double c2 = shp[0][k] * xs[0][1] + shp[1][k] * xs[1][1] + shp[2][k] * xs[2][1];
    double c3 = shp[0][k] * xs[0][2] + shp[1][k] * xs[1][2] + shp[2][k] * xs[2][2];
    shp[0][k] = c1; shp[1][k] = c2; shp[2][k] = c3;
  }
}

/* ======================================================================== */
/* Steel02  (SRC/material/uniaxial/steel/Steel02.cpp)                         */
/* ======================================================================== */
#include <float.h>
typedef struct {
  int kind;
  /* Steel02 parameters (Steel02.h:47) / Concrete02 parameters (Concrete02.cpp:93) */
  double Fy, E0, b, R0, cR1, cR2, a1, a2, a3, a4, sigini;
  double fc, epsc0, fcu, epscu, rat, ft, Ets;
  /* Steel02 history */
  double epsminP, epsmaxP, epsplP, epss0P, sigs0P, epssrP, sigsrP; int konP;
  double epsmin, epsmax, epspl, epss0, sigs0, epsr, sigr; int kon;
  /* Concrete02 history */
  double ecminP, deptP, ecmin, dept;
  /* Steel01 (Steel01.h: fy, E0, b, a1..a4 above; history C* / T*) and ElasticMaterial (Epos, Eneg) */
  double minStrainP, maxStrainP, shiftPP, shiftNP, minStrain, maxStrain, shiftP, shiftN; int loadingP, loading;
  double Epos, Eneg;
  int in_fibre;   /* ElasticMaterial inside a fibre section: the section calls setTrial(), whose tangent at a strain of exactly
                     zero is Epos (ElasticMaterial.cpp:146-160), not getTangent()'s max(Epos, Eneg) (:174-182) */
  /* Concrete01 (Concrete01.h: fpc, epsc0, fpcu, epscu in fc, epsc0, fcu, epscu above; history C* / T*) */
  double endStrainP, unloadSlopeP, endStrain, unloadSlope;     /* (min strain: minStrainP / minStrain) */
  /* ElasticPPMaterial (ElasticPPMaterial.h: fyp, fyn, ezero, E in Epos, plastic strain ep -- updated at commitState) */
  double fyp, fyn, ezero, ep;
  /* common */
  double eP, sigP, epsP, e, sig, eps;
} OrcUni;

/* Steel02::revertToStart, Steel02.cpp:70-104 ; Concrete02 constructor, Concrete02.cpp:93-112 */
static void uni_init(OrcUni* m, int kind, const double* p) {
  memset(m, 0, sizeof *m);
  m->kind = kind;
  if (kind == ORC_UNI_STEEL02) {
    m->Fy = p[0]; m->E0 = p[1]; m->b = p[2]; m->R0 = p[3]; m->cR1 = p[4]; m->cR2 = p[5];
    m->a1 = p[6]; m->a2 = p[7]; m->a3 = p[8]; m->a4 = p[9]; m->sigini = p[10];
    m->eP = m->E0; m->epsP = 0.0; m->sigP = 0.0; m->sig = 0.0; m->eps = 0.0; m->e = m->E0;
    m->konP = 0; m->epsmaxP = m->Fy / m->E0; m->epsminP = -m->epsmaxP;
    m->epsplP = 0.0; m->epss0P = 0.0; m->sigs0P = 0.0; m->epssrP = 0.0; m->sigsrP = 0.0;
    if (m->sigini != 0.0) { m->epsP = m->sigini / m->E0; m->sigP = m->sigini; }
  } else if (kind == ORC_UNI_STEEL01) {
    /* Steel01::Steel01 + revertToStart, Steel01.cpp:40-54, 284-311: p = fy, E0, b, a1, a2, a3, a4 */
    m->Fy = p[0]; m->E0 = p[1]; m->b = p[2]; m->a1 = p[3]; m->a2 = p[4]; m->a3 = p[5]; m->a4 = p[6];
    m->minStrainP = m->maxStrainP = 0.0; m->shiftPP = m->shiftNP = 1.0; m->loadingP = 0;
    m->minStrain = m->maxStrain = 0.0; m->shiftP = m->shiftN = 1.0; m->loading = 0;
    m->epsP = 0.0; m->sigP = 0.0; m->eP = m->E0; m->eps = 0.0; m->sig = 0.0; m->e = m->E0;
  } else if (kind == ORC_UNI_CONCRETE01) {
    /* Concrete01::Concrete01, Concrete01.cpp:89-123: p = fpc, epsc0, fpcu, epscu, made negative */
    m->fc = p[0]; m->epsc0 = p[1]; m->fcu = p[2]; m->epscu = p[3];
    if (m->fc > 0.0) m->fc = -m->fc;
    if (m->epsc0 > 0.0) m->epsc0 = -m->epsc0;
    if (m->fcu > 0.0) m->fcu = -m->fcu;
    if (m->epscu > 0.0) m->epscu = -m->epscu;
    const double Ec0 = 2 * m->fc / m->epsc0;
    m->eP = Ec0; m->unloadSlopeP = Ec0; m->e = Ec0; m->unloadSlope = Ec0;
  } else if (kind == ORC_UNI_ELASTICPP) {
    /* ElasticPPMaterial(tag, E, eyp, eyn, ezero), ElasticPPMaterial.cpp:88-107: p = E, epsyP, epsyN, eps0 */
    double eyp = p[1], eyn = p[2];
    if (eyp < 0) eyp *= -1.;
    if (eyn > 0) eyn *= -1.;
    m->Epos = p[0]; m->fyp = m->Epos * eyp; m->fyn = m->Epos * eyn; m->ezero = p[3]; m->ep = 0.0;
    m->e = m->eP = m->Epos;
  } else if (kind == ORC_UNI_ELASTIC) {
    /* ElasticMaterial(tag, E, eta, Eneg), ElasticMaterial.cpp:96-110: p = E, eta (0 here: no strain rate in this path), Eneg */
    m->Epos = p[0]; m->Eneg = p[2];
    m->e = m->eP = m->Epos > m->Eneg ? m->Epos : m->Eneg;
  } else {
    m->fc = p[0]; m->epsc0 = p[1]; m->fcu = p[2]; m->epscu = p[3]; m->rat = p[4]; m->ft = p[5]; m->Ets = p[6];
    m->ecminP = 0.0; m->deptP = 0.0;
    if (m->fc > 0) m->fc = -m->fc;
    if (m->epsc0 > 0) m->epsc0 = -m->epsc0;
    if (m->fcu > 0) m->fcu = -m->fcu;
    if (m->epscu > 0) m->epscu = -m->epscu;
    m->eP = 2.0 * m->fc / m->epsc0; m->epsP = 0.0; m->sigP = 0.0; m->eps = 0.0; m->sig = 0.0;
    m->e = 2.0 * m->fc / m->epsc0;
  }
}
static double uni_initial_tangent(const OrcUni* m) {
  if (m->kind == ORC_UNI_STEEL01) return m->E0;                                   /* Steel01.h getInitialTangent */
  if (m->kind == ORC_UNI_CONCRETE01) return 2.0 * m->fc / m->epsc0;               /* Concrete01.h getInitialTangent */
  if (m->kind == ORC_UNI_ELASTIC) return m->Epos > m->Eneg ? m->Epos : m->Eneg;   /* ElasticMaterial.cpp:186 */
  if (m->kind == ORC_UNI_ELASTICPP) return m->Epos;                               /* ElasticPPMaterial.h getInitialTangent */
  return m->kind == ORC_UNI_STEEL02 ? m->E0 : 2.0 * m->fc / m->epsc0;   /* Steel02.cpp:107, Concrete02.cpp:161 */
}

/* Steel02::setTrialStrain, Steel02.cpp:113-245 */
static int steel02_set_trial(OrcUni* m, double trialStrain) {
  const double Fy = m->Fy, E0 = m->E0, b = m->b;
  double Esh = b * E0;
  double epsy = Fy / E0;
  if (m->sigini != 0.0) { double epsini = m->sigini / E0; m->eps = trialStrain + epsini; }
  else m->eps = trialStrain;
  double deps = m->eps - m->epsP;
  m->epsmax = m->epsmaxP; m->epsmin = m->epsminP; m->epspl = m->epsplP; m->epss0 = m->epss0P;
  m->sigs0 = m->sigs0P; m->epsr = m->epssrP; m->sigr = m->sigsrP; m->kon = m->konP;
  if (m->kon == 0 || m->kon == 3) {
    if (fabs(deps) < 10.0 * DBL_EPSILON) {
      m->e = E0; m->sig = m->sigini; m->kon = 3;
      return 0;
    } else {
      m->epsmax = epsy; m->epsmin = -epsy;
      if (deps < 0.0) { m->kon = 2; m->epss0 = m->epsmin; m->sigs0 = -Fy; m->epspl = m->epsmin; }
      else { m->kon = 1; m->epss0 = m->epsmax; m->sigs0 = Fy; m->epspl = m->epsmax; }
    }
  }
  if (m->kon == 2 && deps > 0.0) {
    m->kon = 1; m->epsr = m->epsP; m->sigr = m->sigP;
    if (m->epsP < m->epsmin) m->epsmin = m->epsP;
    double d1 = (m->epsmax - m->epsmin) / (2.0 * (m->a4 * epsy));
    double shft = 1.0 + m->a3 * pow(d1, 0.8);
    m->epss0 = (Fy * shft - Esh * epsy * shft - m->sigr + E0 * m->epsr) / (E0 - Esh);
    m->sigs0 = Fy * shft + Esh * (m->epss0 - epsy * shft);
    m->epspl = m->epsmax;
  } else if (m->kon == 1 && deps < 0.0) {
    m->kon = 2; m->epsr = m->epsP; m->sigr = m->sigP;
    if (m->epsP > m->epsmax) m->epsmax = m->epsP;
    double d1 = (m->epsmax - m->epsmin) / (2.0 * (m->a2 * epsy));
    double shft = 1.0 + m->a1 * pow(d1, 0.8);
    m->epss0 = (-Fy * shft + Esh * epsy * shft - m->sigr + E0 * m->epsr) / (E0 - Esh);
    m->sigs0 = -Fy * shft + Esh * (m->epss0 + epsy * shft);
    m->epspl = m->epsmin;
  }
  double xi = fabs((m->epspl - m->epss0) / epsy);
  double R = m->R0 * (1.0 - (m->cR1 * xi) / (m->cR2 + xi));
  double epsrat = (m->eps - m->epsr) / (m->epss0 - m->epsr);
  double dum1 = 1.0 + pow(fabs(epsrat), R);
  double dum2 = pow(dum1, (1 / R));
  m->sig = b * epsrat + (1.0 - b) * epsrat / dum2;
  m->sig = m->sig * (m->sigs0 - m->sigr) + m->sigr;
  m->e = b + (1.0 - b) / (dum1 * dum2);
  m->e = m->e * (m->sigs0 - m->sigr) / (m->epss0 - m->epsr);
  return 0;
}

/* Concrete02::Tens_Envlp / Compr_Envlp, Concrete02.cpp:421-498 */
static void c02_tens(const OrcUni* m, double epsc, double* sigc, double* Ect) {
  double Ec0 = 2.0 * m->fc / m->epsc0;
  double eps0 = m->ft / Ec0;
  double epsu = m->ft * (1.0 / m->Ets + 1.0 / Ec0);
  if (epsc <= eps0) { *sigc = epsc * Ec0; *Ect = Ec0; }
  else if (epsc <= epsu) { *Ect = -m->Ets; *sigc = m->ft - m->Ets * (epsc - eps0); }
  else { *Ect = 1.0e-10; *sigc = 0.0; }
}
static void c02_compr(const OrcUni* m, double epsc, double* sigc, double* Ect) {
  double Ec0 = 2.0 * m->fc / m->epsc0;
  double ratLocal = epsc / m->epsc0;
  if (epsc >= m->epsc0) { *sigc = m->fc * ratLocal * (2.0 - ratLocal); *Ect = Ec0 * (1.0 - ratLocal); }
  else if (epsc > m->epscu) {
    *sigc = (m->fcu - m->fc) * (epsc - m->epsc0) / (m->epscu - m->epsc0) + m->fc;
    *Ect = (m->fcu - m->fc) / (m->epscu - m->epsc0);
  } else { *sigc = m->fcu; *Ect = 1.0e-10; }
}
/* Concrete02::setTrialStrain, Concrete02.cpp:167-270 */
static int concrete02_set_trial(OrcUni* m, double trialStrain) {
  double ec0 = m->fc * 2. / m->epsc0;
  m->ecmin = m->ecminP; m->dept = m->deptP;
  m->eps = trialStrain;
  double deps = m->eps - m->epsP;
  if (fabs(deps) < DBL_EPSILON) return 0;
  if (m->eps < m->ecmin) {
    c02_compr(m, m->eps, &m->sig, &m->e);
    m->ecmin = m->eps;
  } else {
    double epsr = (m->fcu - m->rat * ec0 * m->epscu) / (ec0 * (1.0 - m->rat));
    double sigmr = ec0 * epsr;
    double sigmm, dumy;
    c02_compr(m, m->ecmin, &sigmm, &dumy);
    double er = (sigmm - sigmr) / (m->ecmin - epsr);
    double ept = m->ecmin - sigmm / er;
    if (m->eps <= ept) {
      double sigmin = sigmm + er * (m->eps - m->ecmin);
      double sigmax = er * .5f * (m->eps - ept);
      m->sig = m->sigP + ec0 * deps;
      m->e = ec0;
      if (m->sig <= sigmin) { m->sig = sigmin; m->e = er; }
      if (m->sig >= sigmax) { m->sig = sigmax; m->e = 0.5 * er; }
    } else {
      double epn = ept + m->dept;
      double sicn;
      if (m->eps <= epn) {
        c02_tens(m, m->dept, &sicn, &m->e);
        if (m->dept != 0.0) m->e = sicn / m->dept; else m->e = ec0;
        m->sig = m->e * (m->eps - ept);
      } else {
        double epstmp = m->eps - ept;
        c02_tens(m, epstmp, &m->sig, &m->e);
        m->dept = m->eps - ept;
      }
    }
  }
  return 0;
}
/* Steel01::setTrialStrain + determineTrialState, Steel01.cpp:68-89, 122-196 */
static int steel01_set_trial(OrcUni* m, double strain) {
  m->minStrain = m->minStrainP; m->maxStrain = m->maxStrainP; m->shiftP = m->shiftPP; m->shiftN = m->shiftNP;
  m->loading = m->loadingP; m->eps = m->epsP; m->sig = m->sigP; m->e = m->eP;
  const double dStrain = strain - m->epsP;
  if (!(fabs(dStrain) > DBL_EPSILON)) return 0;
  m->eps = strain;
  const double fy = m->Fy, E0 = m->E0, b = m->b;
  double fyOneMinusB = fy * (1.0 - b);
  double Esh = b * E0;
  double epsy = fy / E0;
  double c1 = Esh * m->eps;
  double c2 = m->shiftN * fyOneMinusB;
  double c3 = m->shiftP * fyOneMinusB;
  double c = m->sigP + E0 * dStrain;
  double c1c3 = c1 + c3;
  if (c1c3 < c) m->sig = c1c3; else m->sig = c;
  double c1c2 = c1 - c2;
  if (c1c2 > m->sig) m->sig = c1c2;
  if (fabs(m->sig - c) < DBL_EPSILON) m->e = E0; else m->e = Esh;
  if (m->loading == 0 && dStrain != 0.0) m->loading = dStrain > 0.0 ? 1 : -1;
  if (m->loading == 1 && dStrain < 0.0) {
    m->loading = -1;
    if (m->epsP > m->maxStrain) m->maxStrain = m->epsP;
    m->shiftN = 1 + m->a1 * pow((m->maxStrain - m->minStrain) / (2.0 * m->a2 * epsy), 0.8);
  }
  if (m->loading == -1 && dStrain > 0.0) {
    m->loading = 1;
    if (m->epsP < m->minStrain) m->minStrain = m->epsP;
    m->shiftP = 1 + m->a3 * pow((m->maxStrain - m->minStrain) / (2.0 * m->a4 * epsy), 0.8);
  }
  return 0;
}
/* ElasticMaterial::setTrialStrain / getStress / getTangent, ElasticMaterial.cpp:137-182 (eta = 0) */
static int elastic_set_trial(OrcUni* m, double strain) {
  m->eps = strain;
  m->sig = strain >= 0.0 ? m->Epos * strain : m->Eneg * strain;
  m->e = strain > 0.0 ? m->Epos : (strain < 0.0 ? m->Eneg : (m->in_fibre ? m->Epos : (m->Epos > m->Eneg ? m->Epos : m->Eneg)));
  return 0;
}
/* Concrete01::setTrialStrain with reload / envelope / unload, Concrete01.cpp:146-206, 313-385 */
static void concrete01_envelope(OrcUni* m) {
  if (m->eps > m->epsc0) {
    double eta = m->eps / m->epsc0;
    m->sig = m->fc * (2 * eta - eta * eta);
    double Ec0 = 2.0 * m->fc / m->epsc0;
    m->e = Ec0 * (1.0 - eta);
  } else if (m->eps > m->epscu) {
    m->e = (m->fc - m->fcu) / (m->epsc0 - m->epscu);
    m->sig = m->fc + m->e * (m->eps - m->epsc0);
  } else { m->sig = m->fcu; m->e = 0.0; }
}
static void concrete01_unload(OrcUni* m) {
  double tempStrain = m->minStrain;
  if (tempStrain < m->epscu) tempStrain = m->epscu;
  double eta = tempStrain / m->epsc0;
  double ratio = 0.707 * (eta - 2.0) + 0.834;
  if (eta < 2.0) ratio = 0.145 * eta * eta + 0.13 * eta;
  m->endStrain = ratio * m->epsc0;
  double temp1 = m->minStrain - m->endStrain;
  double Ec0 = 2.0 * m->fc / m->epsc0;
  double temp2 = m->sig / Ec0;
  if (temp1 > -DBL_EPSILON) m->unloadSlope = Ec0;
  else if (temp1 <= temp2) { m->endStrain = m->minStrain - temp1; m->unloadSlope = m->sig / temp1; }
  else { m->endStrain = m->minStrain - temp2; m->unloadSlope = Ec0; }
}
static void concrete01_reload(OrcUni* m) {
  if (m->eps <= m->minStrain) { m->minStrain = m->eps; concrete01_envelope(m); concrete01_unload(m); }
  else if (m->eps <= m->endStrain) { m->e = m->unloadSlope; m->sig = m->e * (m->eps - m->endStrain); }
  else { m->sig = 0.0; m->e = 0.0; }
}
static int concrete01_set_trial(OrcUni* m, double strain) {
  m->minStrain = m->minStrainP; m->endStrain = m->endStrainP; m->unloadSlope = m->unloadSlopeP;
  m->sig = m->sigP; m->e = m->eP; m->eps = m->epsP;
  double dStrain = strain - m->epsP;
  if (fabs(dStrain) < DBL_EPSILON) return 0;
  m->eps = strain;
  if (m->eps > 0.0) { m->sig = 0; m->e = 0; return 0; }
  m->unloadSlope = m->unloadSlopeP;
  double tempStress = m->sigP + m->unloadSlope * m->eps - m->unloadSlope * m->epsP;
  if (strain < m->epsP) {
    m->minStrain = m->minStrainP; m->endStrain = m->endStrainP;
    concrete01_reload(m);
    if (tempStress > m->sig) { m->sig = tempStress; m->e = m->unloadSlope; }
  } else if (tempStress <= 0.0) { m->sig = tempStress; m->e = m->unloadSlope; }
  else { m->sig = 0.0; m->e = 0.0; }
  return 0;
}
/* ElasticPPMaterial::setTrialStrain, ElasticPPMaterial.cpp:123-167 */
static int elasticpp_set_trial(OrcUni* m, double strain) {
  const double E = m->Epos;
  m->eps = strain;
  double sigtrial = E * (m->eps - m->ezero - m->ep);
  double f;
  if (sigtrial >= 0.0) f = sigtrial - m->fyp; else f = -sigtrial + m->fyn;
  double fYieldSurface = -E * DBL_EPSILON;
  if (f <= fYieldSurface) { m->sig = sigtrial; m->e = E; }
  else { m->sig = sigtrial > 0.0 ? m->fyp : m->fyn; m->e = 0.0; }
  return 0;
}
static int uni_set_trial(OrcUni* m, double strain) {
  if (m->kind == ORC_UNI_ELASTICPP) return elasticpp_set_trial(m, strain);
  if (m->kind == ORC_UNI_CONCRETE01) return concrete01_set_trial(m, strain);
  if (m->kind == ORC_UNI_STEEL01) return steel01_set_trial(m, strain);
  if (m->kind == ORC_UNI_ELASTIC) return elastic_set_trial(m, strain);
  return m->kind == ORC_UNI_STEEL02 ? steel02_set_trial(m, strain) : concrete02_set_trial(m, strain);
}
/* commitState / revertToLastCommit: Steel02.cpp:248-285, Concrete02.cpp:292-320 */
static void uni_commit(OrcUni* m) {
  if (m->kind == ORC_UNI_STEEL02) {
    m->epsminP = m->epsmin; m->epsmaxP = m->epsmax; m->epsplP = m->epspl; m->epss0P = m->epss0;
    m->sigs0P = m->sigs0; m->epssrP = m->epsr; m->sigsrP = m->sigr; m->konP = m->kon;
  } else if (m->kind == ORC_UNI_STEEL01) {   /* Steel01.cpp:244-262 */
    m->minStrainP = m->minStrain; m->maxStrainP = m->maxStrain; m->shiftPP = m->shiftP; m->shiftNP = m->shiftN; m->loadingP = m->loading;
  } else if (m->kind == ORC_UNI_CONCRETE01) {   /* Concrete01.cpp:402-418 */
    m->minStrainP = m->minStrain; m->unloadSlopeP = m->unloadSlope; m->endStrainP = m->endStrain;
  } else if (m->kind == ORC_UNI_CONCRETE02) { m->ecminP = m->ecmin; m->deptP = m->dept; }
  else if (m->kind == ORC_UNI_ELASTICPP) {   /* ElasticPPMaterial::commitState, ElasticPPMaterial.cpp:190-224: the plastic strain moves here */
    const double E = m->Epos;
    double sigtrial = E * (m->eps - m->ezero - m->ep);
    double f;
    if (sigtrial >= 0.0) f = sigtrial - m->fyp; else f = -sigtrial + m->fyn;
    double fYieldSurface = -E * DBL_EPSILON;
    if (f > fYieldSurface) { if (sigtrial > 0.0) m->ep += f / E; else m->ep -= f / E; }
  }
  m->eP = m->e; m->sigP = m->sig; m->epsP = m->eps;
}
static void uni_revert(OrcUni* m) {
  if (m->kind == ORC_UNI_STEEL02) {
    m->epsmin = m->epsminP; m->epsmax = m->epsmaxP; m->epspl = m->epsplP; m->epss0 = m->epss0P;
    m->sigs0 = m->sigs0P; m->epsr = m->epssrP; m->sigr = m->sigsrP; m->kon = m->konP;
  } else if (m->kind == ORC_UNI_STEEL01) {   /* Steel01.cpp:264-281 */
    m->minStrain = m->minStrainP; m->maxStrain = m->maxStrainP; m->shiftP = m->shiftPP; m->shiftN = m->shiftNP; m->loading = m->loadingP;
  } else if (m->kind == ORC_UNI_CONCRETE01) {   /* Concrete01.cpp:420-433 */
    m->minStrain = m->minStrainP; m->endStrain = m->endStrainP; m->unloadSlope = m->unloadSlopeP;
  } else if (m->kind == ORC_UNI_ELASTIC) {   /* ElasticMaterial::revertToLastCommit: the committed strain; stress and tangent follow it */
    elastic_set_trial(m, m->epsP); return;
  } else if (m->kind == ORC_UNI_ELASTICPP) {   /* ElasticPPMaterial::revertToLastCommit: strain, tangent, stress (ep has not moved) */
  } else { m->ecmin = m->ecminP; m->dept = m->deptP; }
  m->e = m->eP; m->sig = m->sigP; m->eps = m->epsP;
}

/* uniaxial strain path; mirrors ref_uni_path in ref_harness.cpp */
int orc_uni_path(int kind, const double* p, int n, const double* strains, const int* commit,
                 double* stress, double* tangent) {
  OrcUni m; uni_init(&m, kind, p);
  for (int s = 0; s < n; s++) {
    if (uni_set_trial(&m, strains[s]) < 0) return -1;
    stress[s] = m.sig; tangent[s] = m.e;
    if (commit[s]) uni_commit(&m);
  }
  return 0;
}

/* ======================================================================== */
/* FiberSection2d (SRC/material/section/FiberSection2d.cpp)                   */
/* ======================================================================== */
typedef struct {
  int nf; double* y; double* A; OrcUni* mat; double yBar;
  double e[2], s[2], k[4];     /* kData column-major 2x2: k[0]=k00 k[1]=k10 k[2]=k01 k[3]=k11 */
  int agg;                     /* section Aggregator of two uniaxial materials, codes P and Mz (SectionAggregator.cpp:119):
                                  mat[0] takes the axial strain, mat[1] the curvature; no fibres */
} OrcSec;

/* FiberSection2d::setTrialSectionDeformation, FiberSection2d.cpp:225-262 */
static int sec_set_trial(OrcSec* S, const double* d) {
  S->e[0] = d[0]; S->e[1] = d[1];
  S->k[0] = S->k[1] = S->k[2] = S->k[3] = 0.0; S->s[0] = S->s[1] = 0.0;
  const double d0 = d[0], d1 = d[1];
  int res = 0;
  if (S->agg) {   /* SectionAggregator::setTrialSectionDeformation / getSectionTangent / getStressResultant, :316, :365, :483 */
    res += uni_set_trial(&S->mat[0], d0);
    res += uni_set_trial(&S->mat[1], d1);
    S->k[0] = S->mat[0].e; S->k[3] = S->mat[1].e;
    S->s[0] = S->mat[0].sig; S->s[1] = S->mat[1].sig;
    return res;
  }
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, A = S->A[i];
    double strain = d0 - y * d1;
    res += uni_set_trial(&S->mat[i], strain);
    double tangent = S->mat[i].e, stress = S->mat[i].sig;
    double ks0 = tangent * A;
    double ks1 = ks0 * -y;
    S->k[0] += ks0; S->k[1] += ks1; S->k[3] += ks1 * -y;
    double fs0 = stress * A;
    S->s[0] += fs0; S->s[1] += fs0 * -y;
  }
  S->k[2] = S->k[1];
  return res;
}
/* FiberSection2d::revertToLastCommit, FiberSection2d.cpp:376-408 (note sData is ASSIGNED, not summed) */
static void sec_revert(OrcSec* S) {
  S->k[0] = S->k[1] = S->k[2] = S->k[3] = 0.0; S->s[0] = S->s[1] = 0.0;
  if (S->agg) {   /* SectionAggregator::revertToLastCommit, :569: the materials only */
    uni_revert(&S->mat[0]); uni_revert(&S->mat[1]);
    S->k[0] = S->mat[0].e; S->k[3] = S->mat[1].e; S->s[0] = S->mat[0].sig; S->s[1] = S->mat[1].sig;
    return;
  }
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, A = S->A[i];
    uni_revert(&S->mat[i]);
    double ks0 = S->mat[i].e * A, ks1 = ks0 * -y;
    S->k[0] += ks0; S->k[1] += ks1; S->k[3] += ks1 * -y;
    double fs0 = S->mat[i].sig * A;
    S->s[0] = fs0; S->s[1] = fs0 * -y;
  }
  S->k[2] = S->k[1];
}
/* SectionForceDeformation::getSectionFlexibility -> Matrix::Invert -> cmx_inv2
 * (SectionForceDeformation.cpp, matrix/routines/invGL2.c); k, f column-major */
static void inv2(const double* a, double* ainv) {
  const double det = a[0] * a[3] - a[2] * a[1];
  ainv[0] = a[3] / det; ainv[1] = -a[1] / det; ainv[2] = -a[2] / det; ainv[3] = a[0] / det;
}
/* matrix/routines/invGL3.c (column-major 3x3) */
static void inv3(const double* a, double* ainv) {
  const double* A = a - 4; double* I = ainv - 4;
  const double det = A[4]*A[8]*A[12] - A[4]*A[11]*A[9] - A[7]*A[5]*A[12] + A[7]*A[11]*A[6] + A[10]*A[5]*A[9] - A[10]*A[8]*A[6];
  double c[9];
  c[0] =  A[8]*A[12] - A[11]*A[9];  c[3] = -(A[5]*A[12] - A[11]*A[6]); c[6] =  A[5]*A[9] - A[8]*A[6];
  c[1] = -(A[7]*A[12] - A[10]*A[9]); c[4] =  A[4]*A[12] - A[10]*A[6];  c[7] = -(A[4]*A[9] - A[7]*A[6]);
  c[2] =  A[7]*A[11] - A[10]*A[8];  c[5] = -(A[4]*A[11] - A[10]*A[5]); c[8] =  A[4]*A[8] - A[7]*A[5];
  for (int i = 1; i <= 3; ++i) for (int j = 1; j <= 3; ++j) I[j + i * 3] = c[i + j * 3 - 4] / det;
}
/* the section's getSectionFlexibility: the inverse of the tangent (fibre section), or SectionAggregator's own
 * (SectionAggregator.cpp:419-450): 1/k on the diagonal, 1e14 for a zero tangent */
static void sec_flex(const OrcSec* S, double* f) {
  if (S->agg) {
    f[1] = f[2] = 0.0;
    f[0] = S->k[0] == 0.0 ? 1.e14 : 1 / S->k[0];
    f[3] = S->k[3] == 0.0 ? 1.e14 : 1 / S->k[3];
    return;
  }
  inv2(S->k, f);
}
static void sec_initial_flex(const OrcSec* S, double* f) {   /* FiberSection2d::getInitialTangent + Invert */
  double k[4] = {0, 0, 0, 0};
  if (S->agg) {   /* SectionAggregator::getInitialFlexibility, :454-479 */
    f[1] = f[2] = 0.0; f[0] = 1.0 / uni_initial_tangent(&S->mat[0]); f[3] = 1.0 / uni_initial_tangent(&S->mat[1]);
    return;
  }
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, A = S->A[i];
    double ks0 = uni_initial_tangent(&S->mat[i]) * A, ks1 = ks0 * -y;
    k[0] += ks0; k[1] += ks1; k[3] += ks1 * -y;
  }
  k[2] = k[1];
  inv2(k, f);
}

/* ======================================================================== */
/* ForceBeamColumn2d (SRC/element/Frame/Other/Force/ForceBeamColumn2d.cpp)    */
/* with LinearCrdTransf2d and LobattoBeamIntegration                          */
/* ======================================================================== */
#define ORC_MAXSEC 10
typedef struct {
  int nip, maxIters; double tol;
  OrcSec sec[ORC_MAXSEC];
  double L, cosTheta, sinTheta;
  int initialFlag;
  double kv[9], Se[3], kvcommit[9], Secommit[3];       /* kv column-major 3x3 */
  double fs[ORC_MAXSEC][4], vs[ORC_MAXSEC][2], Ssr[ORC_MAXSEC][2], vscommit[ORC_MAXSEC][2];
  /* `eleLoad -beamUniform wy wa` of the Linear pattern: w = {wy, -, wa}; numEleLoads = 1 once Domain::applyLoad ran
   * (LoadPattern::applyLoad -> ElementalLoad::applyLoad -> ForceBeamColumn2d::addLoad(load, loadFactor)) */
  int has_load, numEleLoads; double w[3], loadFactor;
  /* `eleLoad -beamPoint Py xL [N]` (Beam2dPointLoad): pt = {Py, -, N, aOverL} */
  int has_point; double pt[4];
  /* `eleLoad -beamUniform` over part of the element (Beam2dPartialUniformLoad): pq = wTrans_a, wTrans_b, wAxial_a,
   * wAxial_b, aOverL, bOverL (getData's order) */
  int has_partial; double pq[6];
  /* beam integration other than Lobatto: the section locations and weights the element's BeamIntegration object returns
   * (getSectionLocations / getSectionWeights), handed over by the caller (orc_set_beam_integration) */
  int user_rule; double rxi[ORC_MAXSEC], rwt[ORC_MAXSEC];
  /* geomTransf PDelta (PDeltaCrdTransf2d.cpp): ul14 is recomputed from the nodes' trial displacements whenever the
   * element asks for its tangent or resisting force (ForceBeamColumn2d.cpp:402,526 call crdTransf->update()) */
  int pdelta; const double* utrial; int n0, n1;
  /* rigid joint offsets (geomTransf ... -jntOffset dXi dYi dXj dYj; LinearCrdTransf2d.cpp / PDeltaCrdTransf2d.cpp nodeIOffset,
   * nodeJOffset): the element ends sit at node + offset and follow the node rigidly, u_end = u + theta x offset */
  int has_off; double off[4];
  /* geomTransf Corotational (CorotCrdTransf2d.cpp, no joint offsets): the basic deformations of the last
   * crdTransf->update() (ub; the next update's ubpr) and of the last commit (ubcommit) */
  int corot; double cub[3], cubc[3];
} OrcBeam;

/* quadrature/Frame/LobattoBeamIntegration.cpp: getSectionLocations / getSectionWeights */
static int lobatto(int n, double* xi, double* wt) {
  static const double X[11][10] = {{0},{0},{-1.0,1.0},{-1.0,0.0,1.0},{-1.0,-0.44721360,0.44721360,1.0},
    {-1.0,-0.65465367,0.0,0.65465367,1.0},{-1.0,-0.7650553239,-0.2852315164,0.2852315164,0.7650553239,1.0},
    {-1.0,-0.8302238962,-0.4688487934,0.0,0.4688487934,0.8302238962,1.0},
    {-1.0,-0.8717401485,-0.5917001814,-0.2092992179,0.2092992179,0.5917001814,0.8717401485,1.0},
    {-1.0,-0.8997579954,-0.6771862795,-0.3631174638,0.0,0.3631174638,0.6771862795,0.8997579954,1.0},
    {-1.0,-0.9195339082,-0.7387738651,-0.4779249498,-0.1652789577,0.1652789577,0.4779249498,0.7387738651,0.9195339082,1.0}};
  static const double W[11][10] = {{0},{0},{1.0,1.0},{0.333333333333333,1.333333333333333,0.333333333333333},
    {0.166666666666667,0.833333333333333,0.833333333333333,0.166666666666667},
    {0.1,0.5444444444,0.7111111111,0.5444444444,0.1},
    {0.06666666667,0.3784749562,0.5548583770,0.5548583770,0.3784749562,0.06666666667},
    {0.04761904762,0.2768260473,0.4317453812,0.4876190476,0.4317453812,0.2768260473,0.04761904762},
    {0.03571428571,0.2107042271,0.3411226924,0.4124587946,0.4124587946,0.3411226924,0.2107042271,0.03571428571},
    {0.02777777778,0.1654953615,0.2745387125,0.3464285109,0.3715192743,0.3464285109,0.2745387125,0.1654953615,0.02777777778},
    {0.02222222222,0.1333059908,0.2248893421,0.2920426836,0.3275397611,0.3275397611,0.2920426836,0.2248893421,0.1333059908,0.02222222222}};
  if (n < 2 || n > 10) return -1;
  for (int i = 0; i < n; i++) { xi[i] = 0.5 * (X[n][i] + 1.0); wt[i] = W[n][i] * 0.5; }
  return 0;
}

/* LinearCrdTransf2d::getBasicTrialDisp / getBasicIncrDeltaDisp (no offsets), LinearCrdTransf2d.cpp */
static void crd2d_basic(const OrcBeam* b, const double* ug, double* ub) {
  double oneOverL = 1.0 / b->L;
  double sl = b->sinTheta * oneOverL, cl = b->cosTheta * oneOverL;
  ub[0] = -b->cosTheta * ug[0] - b->sinTheta * ug[1] + b->cosTheta * ug[3] + b->sinTheta * ug[4];
  ub[1] = -sl * ug[0] + cl * ug[1] + ug[2] + sl * ug[3] - cl * ug[4];
  ub[2] = ub[1] + ug[5] - ug[2];
}

/* Beam2dPartialUniformLoad: reactions (computeReactions, ForceBeamColumn2d.cpp:426-443) and section forces at x
 * (computeSectionForces, :1073-1137) */
static void beam2_partial_p0(const OrcBeam* bm, double* p0) {
  if (!bm->has_partial) return;
  const double lf = bm->loadFactor, L = bm->L;
  double waa = bm->pq[2] * lf, wab = bm->pq[3] * lf, wya = bm->pq[0] * lf, wyb = bm->pq[1] * lf;
  double a = bm->pq[4] * L, b = bm->pq[5] * L;
  p0[0] -= waa * (b - a) + 0.5 * (wab - waa) * (b - a);
  double Fy = wya * (b - a);
  double c = a + 0.5 * (b - a);
  p0[1] -= Fy * (1 - c / L);
  p0[2] -= Fy * c / L;
  Fy = 0.5 * (wyb - wya) * (b - a);
  c = a + 2.0 / 3.0 * (b - a);
  p0[1] -= Fy * (1 - c / L);
  p0[2] -= Fy * c / L;
}
static void beam2_partial_sp(const OrcBeam* bm, double x, double* Ss) {
  if (!bm->has_partial) return;
  const double lf = bm->loadFactor, L = bm->L;
  double waa = bm->pq[2] * lf, wab = bm->pq[3] * lf, wya = bm->pq[0] * lf, wyb = bm->pq[1] * lf;
  double a = bm->pq[4] * L, b = bm->pq[5] * L;
  double Fa = waa * (b - a) + 0.5 * (wab - waa) * (b - a);
  double Fy = wya * (b - a);
  double c = a + 0.5 * (b - a);
  double VI = Fy * (1 - c / L), VJ = Fy * c / L;
  Fy = 0.5 * (wyb - wya) * (b - a);
  c = a + 2.0 / 3.0 * (b - a);
  VI += Fy * (1 - c / L); VJ += Fy * c / L;
  if (x <= a) { Ss[0] += Fa; Ss[1] -= VI * x; }
  else if (x >= b) Ss[1] += VJ * (x - L);
  else {
    double wx = wya + (wyb - wya) / (b - a) * (x - a);
    Ss[0] += Fa - waa * (x - a) - 0.5 * (wab - waa) / (b - a) * (x - a) * (x - a);
    Ss[1] += -VI * x + wya * (x - a) * 0.5 * (x - a) + 0.5 * (wx - wya) * (x - a) * (x - a) / 3.0;
  }
}

/* CorotCrdTransf2d::update (CorotCrdTransf2d.cpp:179-231, no offsets): local end displacements, the deformed chord
 * (compElemtLengthAndOrientWRTLocalSystem, :272-299) and the basic deformations with the rigid-body rotation alpha
 * taken out (transfLocalDisplsToBasic, :344-355).  cg = Lx, Ly, Ln, cosAlpha, sinAlpha */
static int corot2d_geom(const OrcBeam* b, const double* ug, double* cg, double* ub) {
  const double cosTheta = b->cosTheta, sinTheta = b->sinTheta;
  double ul[6];
  ul[0] = cosTheta * ug[0] + sinTheta * ug[1];
  ul[1] = cosTheta * ug[1] - sinTheta * ug[0];
  ul[2] = ug[2];
  ul[3] = cosTheta * ug[3] + sinTheta * ug[4];
  ul[4] = cosTheta * ug[4] - sinTheta * ug[3];
  ul[5] = ug[5];
  const double dulx = ul[3] - ul[0], duly = ul[4] - ul[1];
  const double Lx = b->L + dulx, Ly = duly;
  const double Ln = sqrt(Lx * Lx + Ly * Ly);
  if (Ln == 0.0) return -2;
  const double cosAlpha = Lx / Ln, sinAlpha = Ly / Ln;
  cg[0] = Lx; cg[1] = Ly; cg[2] = Ln; cg[3] = cosAlpha; cg[4] = sinAlpha;
  if (ub) {
    const double alpha = atan2(sinAlpha, cosAlpha);
    ub[0] = Ln - b->L; ub[1] = ul[2] - alpha; ub[2] = ul[5] - alpha;
  }
  return 0;
}

/* ForceBeamColumn2d::update, ForceBeamColumn2d.cpp:559-933 (no element loads) */
static void beam_end_disp(const OrcBeam* b, double* ug);
static int beam_update(OrcBeam* b, const double* ug_, const double* dug_) {
  double v[3], dv[3], vin[3];
  double ug[6], dug[6];
  memcpy(ug, ug_, sizeof ug); memcpy(dug, dug_, sizeof dug);
  beam_end_disp(b, ug); beam_end_disp(b, dug);
  if (b->corot) {   /* CorotCrdTransf2d::update (:179-231), getBasicIncrDeltaDisp = ub - ubpr (:364-371) */
    double cg[5];
    if (corot2d_geom(b, ug, cg, v) < 0) return -1;
    for (int i = 0; i < 3; i++) { dv[i] = v[i] - b->cub[i]; b->cub[i] = v[i]; }
  } else {
  crd2d_basic(b, ug, v);
  crd2d_basic(b, dug, dv);
  }
  double nrm = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
  if (b->initialFlag != 0 && nrm <= DBL_EPSILON && b->numEleLoads == 0) return 0;
  for (int i = 0; i < 3; i++) vin[i] = v[i] - dv[i];
  const double L = b->L, oneOverL = 1.0 / L;
  double xi[ORC_MAXSEC], wt[ORC_MAXSEC];
  if (b->user_rule) { memcpy(xi, b->rxi, sizeof(double) * b->nip); memcpy(wt, b->rwt, sizeof(double) * b->nip); } else lobatto(b->nip, xi, wt);
  double vr[3], f[9], dSe[3], SeTrial[3], kvTrial[9], dvTrial[3], dvToDo[3];
  double vsSub[ORC_MAXSEC][2], fsSub[ORC_MAXSEC][4], SsrSub[ORC_MAXSEC][2];
  int numSubdivide = 1, converged = 0;
  for (int i = 0; i < 3; i++) { dvToDo[i] = dv[i]; dvTrial[i] = dvToDo[i]; }
  const double factor = 10;
  const int maxSubdivisions = 4;
  (void)oneOverL;
  while (!converged && numSubdivide <= maxSubdivisions) {
    for (int l = 0; l < 3; l++) {
      memcpy(SeTrial, b->Se, sizeof SeTrial); memcpy(kvTrial, b->kv, sizeof kvTrial);
      for (int i = 0; i < b->nip; i++) {
        memcpy(vsSub[i], b->vs[i], sizeof vsSub[i]); memcpy(fsSub[i], b->fs[i], sizeof fsSub[i]);
        memcpy(SsrSub[i], b->Ssr[i], sizeof SsrSub[i]);
      }
      /* dSe = kv * dv (Matrix::addMatrixVector: column by column) */
      for (int i = 0; i < 3; i++) dSe[i] = 0.0;
      for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) dSe[i] += kvTrial[i + 3 * j] * dvTrial[j];
      for (int i = 0; i < 3; i++) SeTrial[i] += dSe[i];
      int numIters = b->maxIters;
      if (l == 1) numIters = 10 * b->maxIters;
      for (int j = 0; j < numIters; j++) {
        for (int i = 0; i < 9; i++) f[i] = 0.0;
        vr[0] = vr[1] = vr[2] = 0.0;
        for (int i = 0; i < b->nip; i++) {
          OrcSec* S = &b->sec[i];
          double Ss[2], dSs[2], dvs[2], fb[6];   /* fb (order x 3) column-major */
          double xL = xi[i], xL1 = xL - 1.0, wtL = wt[i] * L;
          Ss[0] = SeTrial[0];
          Ss[1] = xL1 * SeTrial[1] + xL * SeTrial[2];
          if (b->numEleLoads > 0) {   /* computeSectionForces, ForceBeamColumn2d.cpp:1034-1070 (Beam2dUniformLoad) */
            double x = xi[i] * L;
            double wa = b->w[2] * b->loadFactor, wy = b->w[0] * b->loadFactor;
            Ss[0] += wa * (L - x);
            Ss[1] += wy * 0.5 * x * (x - L);
            if (b->has_point) {       /* Beam2dPointLoad, ForceBeamColumn2d.cpp:1138-1181 */
              double P = b->pt[0] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
              double a = aOverL * L;
              double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
              if (x <= a) { Ss[0] += N; Ss[1] -= x * V1; }
              else Ss[1] -= (L - x) * V2;
            }
            beam2_partial_sp(b, x, Ss);
          }
          dSs[0] = Ss[0] - SsrSub[i][0]; dSs[1] = Ss[1] - SsrSub[i][1];
          const double* fuse;
          double fs0[4];
          if (l == 0) fuse = fsSub[i];
          else if (l == 2) { if (j == 0) { sec_initial_flex(S, fs0); fuse = fs0; } else fuse = fsSub[i]; }
          else { sec_initial_flex(S, fs0); fuse = fs0; }
          /* dvs = fs * dSs */
          dvs[0] = 0.0; dvs[1] = 0.0;
          for (int c = 0; c < 2; c++) for (int r = 0; r < 2; r++) dvs[r] += fuse[r + 2 * c] * dSs[c];
          if (b->initialFlag != 0) { vsSub[i][0] += dvs[0]; vsSub[i][1] += dvs[1]; }
          if (sec_set_trial(S, vsSub[i]) < 0) return -1;
          SsrSub[i][0] = S->s[0]; SsrSub[i][1] = S->s[1];
          sec_flex(S, fsSub[i]);
          dSs[0] = Ss[0] - SsrSub[i][0]; dSs[1] = Ss[1] - SsrSub[i][1];
          dvs[0] = 0.0; dvs[1] = 0.0;
          for (int c = 0; c < 2; c++) for (int r = 0; r < 2; r++) dvs[r] += fsSub[i][r + 2 * c] * dSs[c];
          /* fb = fs * b * wtL ; code = {P, MZ} */
          for (int q = 0; q < 6; q++) fb[q] = 0.0;
          const double* fSec = fsSub[i];
          for (int jj = 0; jj < 2; jj++) fb[jj + 2 * 0] += fSec[jj + 2 * 0] * wtL;
          for (int jj = 0; jj < 2; jj++) { double tmp = fSec[jj + 2 * 1] * wtL; fb[jj + 2 * 1] += xL1 * tmp; fb[jj + 2 * 2] += xL * tmp; }
          /* f += b^T fb */
          for (int jj = 0; jj < 3; jj++) f[0 + 3 * jj] += fb[0 + 2 * jj];
          for (int jj = 0; jj < 3; jj++) { double tmp = fb[1 + 2 * jj]; f[1 + 3 * jj] += xL1 * tmp; f[2 + 3 * jj] += xL * tmp; }
          /* vr += b^T (vs + dvs) wtL */
          dvs[0] += vsSub[i][0]; dvs[1] += vsSub[i][1];
          { double dei = dvs[0] * wtL; vr[0] += dei; }
          { double dei = dvs[1] * wtL; vr[1] += xL1 * dei; vr[2] += xL * dei; }
        }
        inv3(f, kvTrial);
        for (int i = 0; i < 3; i++) { dv[i] = vin[i]; dv[i] += dvTrial[i]; dv[i] -= vr[i]; }
        for (int i = 0; i < 3; i++) dSe[i] = 0.0;
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) dSe[r] += kvTrial[r + 3 * c] * dv[c];
        double dW = 0.0;
        for (int i = 0; i < 3; i++) dW += dv[i] * dSe[i];
        for (int i = 0; i < 3; i++) SeTrial[i] += dSe[i];
        if (fabs(dW) < b->tol) {
          for (int i = 0; i < 3; i++) { dvToDo[i] -= dvTrial[i]; vin[i] += dvTrial[i]; }
          double nn = sqrt(dvToDo[0] * dvToDo[0] + dvToDo[1] * dvToDo[1] + dvToDo[2] * dvToDo[2]);
          if (nn <= DBL_EPSILON) converged = 1;
          else { for (int i = 0; i < 3; i++) dvTrial[i] = dvToDo[i]; numSubdivide = 1; }
          memcpy(b->kv, kvTrial, sizeof kvTrial); memcpy(b->Se, SeTrial, sizeof SeTrial);
          for (int k = 0; k < b->nip; k++) {
            memcpy(b->vs[k], vsSub[k], sizeof vsSub[k]); memcpy(b->fs[k], fsSub[k], sizeof fsSub[k]);
            memcpy(b->Ssr[k], SsrSub[k], sizeof SsrSub[k]);
          }
          j = numIters + 1; l = 3;
        } else {
          if (j == (numIters - 1) && (l == 2)) { for (int i = 0; i < 3; i++) dvTrial[i] /= factor; numSubdivide++; }
        }
      }
    }
  }
  if (!converged) return -1;
  b->initialFlag = 1;
  return 0;
}

/* PDeltaCrdTransf2d::update / getGlobalStiffMatrix / getGlobalResistingForce (no offsets), PDeltaCrdTransf2d.cpp:349-384,
 * 566-745, 507-564; K row-major 6x6 */
/* displacements of the element ends from those of the nodes (joint offsets): u_end = u + theta x offset */
static void beam_end_disp(const OrcBeam* b, double* ug) {
  if (!b->has_off) return;
  ug[0] += -ug[2] * b->off[1]; ug[1] += ug[2] * b->off[0];
  ug[3] += -ug[5] * b->off[3]; ug[4] += ug[5] * b->off[2];
}
static void beam_form_pdelta(const OrcBeam* b, double* K, double* R) {
  const double cosTheta = b->cosTheta, sinTheta = b->sinTheta, oneOverL = 1.0 / b->L;
  double ue[6];
  for (int j = 0; j < 3; j++) { ue[j] = b->utrial[3 * b->n0 + j]; ue[3 + j] = b->utrial[3 * b->n1 + j]; }
  beam_end_disp(b, ue);
  const double* uI = ue; const double* uJ = ue + 3;
  const double ul1 = -sinTheta * uI[0] + cosTheta * uI[1];
  const double ul4 = -sinTheta * uJ[0] + cosTheta * uJ[1];
  const double ul14 = ul1 - ul4;
  if (K) {
    const double* kb = b->kv;
    double kb00 = kb[0], kb10 = kb[1], kb20 = kb[2], kb01 = kb[3], kb11 = kb[4], kb21 = kb[5], kb02 = kb[6], kb12 = kb[7], kb22 = kb[8];
    double kl[6][6], tmp[6][6];
    kl[0][0] = kb00; kl[1][0] = -oneOverL * (kb10 + kb20); kl[2][0] = -kb10; kl[3][0] = -kb00; kl[4][0] = -kl[1][0]; kl[5][0] = -kb20;
    kl[0][1] = -oneOverL * (kb01 + kb02); kl[1][1] = oneOverL * oneOverL * (kb11 + kb12 + kb21 + kb22); kl[2][1] = oneOverL * (kb11 + kb12);
    kl[3][1] = -kl[0][1]; kl[4][1] = -kl[1][1]; kl[5][1] = oneOverL * (kb21 + kb22);
    kl[0][2] = -kb01; kl[1][2] = oneOverL * (kb11 + kb21); kl[2][2] = kb11; kl[3][2] = kb01; kl[4][2] = -kl[1][2]; kl[5][2] = kb21;
    for (int i = 0; i < 6; i++) { kl[i][3] = -kl[i][0]; kl[i][4] = -kl[i][1]; }
    kl[0][5] = -kb02; kl[1][5] = oneOverL * (kb12 + kb22); kl[2][5] = kb12; kl[3][5] = kb02; kl[4][5] = -kl[1][5]; kl[5][5] = kb22;
    const double NoverL = b->Se[0] * oneOverL;          /* geometric stiffness, :628-633 */
    kl[1][1] += NoverL; kl[4][4] += NoverL; kl[1][4] -= NoverL; kl[4][1] -= NoverL;
    for (int i = 0; i < 6; i++) {                       /* kl * T, :650-700 */
      tmp[i][0] = kl[i][0] * cosTheta - kl[i][1] * sinTheta;
      tmp[i][1] = kl[i][0] * sinTheta + kl[i][1] * cosTheta;
      tmp[i][2] = kl[i][2];
      tmp[i][3] = kl[i][3] * cosTheta - kl[i][4] * sinTheta;
      tmp[i][4] = kl[i][3] * sinTheta + kl[i][4] * cosTheta;
      tmp[i][5] = kl[i][5];
    }
    for (int j = 0; j < 6; j++) {                       /* T^T * (kl * T), :702-745 */
      K[0 * 6 + j] = cosTheta * tmp[0][j] - sinTheta * tmp[1][j];
      K[1 * 6 + j] = sinTheta * tmp[0][j] + cosTheta * tmp[1][j];
      K[2 * 6 + j] = tmp[2][j];
      K[3 * 6 + j] = cosTheta * tmp[3][j] - sinTheta * tmp[4][j];
      K[4 * 6 + j] = sinTheta * tmp[3][j] + cosTheta * tmp[4][j];
      K[5 * 6 + j] = tmp[5][j];
    }
  }
  double q0 = b->Se[0], q1 = b->Se[1], q2 = b->Se[2];
  double V = oneOverL * (q1 + q2);
  double pl[6] = { -q0, V, q1, q0, -V, q2 };
  double p0[3] = {0.0, 0.0, 0.0};
  if (b->numEleLoads > 0) {
    double wa = b->w[2] * b->loadFactor, wy = b->w[0] * b->loadFactor;
    p0[0] -= wa * b->L;
    double Vr = 0.5 * wy * b->L;
    p0[1] -= Vr; p0[2] -= Vr;
    if (b->has_point) {               /* Beam2dPointLoad, ForceBeamColumn2d.cpp:442-455 */
      double P = b->pt[0] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
      double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
      p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
    }
    beam2_partial_p0(b, p0);
  }
  pl[0] += p0[0]; pl[1] += p0[1]; pl[4] += p0[2];
  double NoverL = ul14 * q0 * oneOverL;                 /* leaning-column effect, :532-535 */
  pl[1] += NoverL; pl[4] -= NoverL;
  R[0] = cosTheta * pl[0] - sinTheta * pl[1];
  R[1] = sinTheta * pl[0] + cosTheta * pl[1];
  R[3] = cosTheta * pl[3] - sinTheta * pl[4];
  R[4] = sinTheta * pl[3] + cosTheta * pl[4];
  R[2] = pl[2]; R[5] = pl[5];
}
/* CorotCrdTransf2d::getGlobalStiffMatrix (:525-701) / getGlobalResistingForce (:483-522), no offsets; the 2D element
 * refreshes the transformation from the nodes' trial displacements before either (ForceBeamColumn2d.cpp:402,526).
 * K row-major 6x6 */
static void beam_form_corot(const OrcBeam* b, double* K, double* R) {
  const double cosTheta = b->cosTheta, sinTheta = b->sinTheta;
  double ue[6], cg[5];
  for (int j = 0; j < 3; j++) { ue[j] = b->utrial[3 * b->n0 + j]; ue[3 + j] = b->utrial[3 * b->n1 + j]; }
  corot2d_geom(b, ue, cg, NULL);
  const double Ln = cg[2], cosAlpha = cg[3], sinAlpha = cg[4];
  double Tbl[3][6];        /* compTransfMatrixBasicLocal, :315-341 */
  Tbl[0][0] = -cosAlpha; Tbl[1][0] = -sinAlpha / Ln; Tbl[2][0] = -sinAlpha / Ln;
  Tbl[0][1] = -sinAlpha; Tbl[1][1] = cosAlpha / Ln; Tbl[2][1] = cosAlpha / Ln;
  Tbl[0][2] = 0; Tbl[1][2] = 1; Tbl[2][2] = 0;
  Tbl[0][3] = cosAlpha; Tbl[1][3] = sinAlpha / Ln; Tbl[2][3] = sinAlpha / Ln;
  Tbl[0][4] = sinAlpha; Tbl[1][4] = -cosAlpha / Ln; Tbl[2][4] = -cosAlpha / Ln;
  Tbl[0][5] = 0; Tbl[1][5] = 0; Tbl[2][5] = 1;
  const double* pb = b->Se;
  if (K) {
    double kl[6][6], tk[3][6];
    /* kl = Tbl^T kb Tbl (Matrix::addMatrixTripleProduct); kv column-major 3x3 */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 6; j++) {
      double t = 0.0;
      for (int k = 0; k < 3; k++) t += b->kv[i + 3 * k] * Tbl[k][j];
      tk[i][j] = t;
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
      double t = 0.0;
      for (int k = 0; k < 3; k++) t += Tbl[k][i] * tk[k][j];
      kl[i][j] = t;
    }
    /* getGeomStiffMatrix, :908-953 */
    const double s2 = sinAlpha * sinAlpha, c2 = cosAlpha * cosAlpha, cs = sinAlpha * cosAlpha;
    double kg0[6][6], kg12[6][6];
    memset(kg0, 0, sizeof kg0); memset(kg12, 0, sizeof kg12);
    kg0[0][0] = kg0[3][3] = s2; kg0[0][1] = kg0[3][4] = -cs; kg0[1][0] = kg0[4][3] = -cs; kg0[1][1] = kg0[4][4] = c2;
    kg0[0][3] = kg0[3][0] = -s2; kg0[0][4] = kg0[3][1] = cs; kg0[1][3] = kg0[4][0] = cs; kg0[1][4] = kg0[4][1] = -c2;
    kg12[0][0] = kg12[3][3] = -2 * cs; kg12[0][1] = kg12[3][4] = c2 - s2; kg12[1][0] = kg12[4][3] = c2 - s2; kg12[1][1] = kg12[4][4] = 2 * cs;
    kg12[0][3] = kg12[3][0] = 2 * cs; kg12[0][4] = kg12[3][1] = -c2 + s2; kg12[1][3] = kg12[4][0] = -c2 + s2; kg12[1][4] = kg12[4][1] = -2 * cs;
    const double f0 = pb[0] / Ln, f12 = (pb[1] + pb[2]) / (Ln * Ln);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) kl[i][j] += kg0[i][j] * f0 + kg12[i][j] * f12;
    /* kg = Tlg^T kl Tlg, block by block, :543-644 */
    const double S2 = sinTheta * sinTheta, C2 = cosTheta * cosTheta, CS = sinTheta * cosTheta;
    for (int bi = 0; bi < 2; bi++) for (int bj = 0; bj < 2; bj++) {
      const int r = 3 * bi, c = 3 * bj;
      const double k11 = kl[r][c], k12 = kl[r][c + 1], k13 = kl[r][c + 2], k21 = kl[r + 1][c], k22 = kl[r + 1][c + 1], k23 = kl[r + 1][c + 2],
                   k31 = kl[r + 2][c], k32 = kl[r + 2][c + 1], k33 = kl[r + 2][c + 2];
      K[(r + 0) * 6 + c + 0] = C2 * k11 + S2 * k22 - CS * (k21 + k12);
      K[(r + 1) * 6 + c + 0] = C2 * k21 - S2 * k12 + CS * (k11 - k22);
      K[(r + 2) * 6 + c + 0] = cosTheta * k31 - sinTheta * k32;
      K[(r + 0) * 6 + c + 1] = C2 * k12 - S2 * k21 + CS * (k11 - k22);
      K[(r + 1) * 6 + c + 1] = C2 * k22 + S2 * k11 + CS * (k21 + k12);
      K[(r + 2) * 6 + c + 1] = sinTheta * k31 + cosTheta * k32;
      K[(r + 0) * 6 + c + 2] = cosTheta * k13 - sinTheta * k23;
      K[(r + 1) * 6 + c + 2] = sinTheta * k13 + cosTheta * k23;
      K[(r + 2) * 6 + c + 2] = k33;
    }
  }
  double pl[6];
  for (int j = 0; j < 6; j++) pl[j] = Tbl[0][j] * pb[0] + Tbl[1][j] * pb[1] + Tbl[2][j] * pb[2];   /* pl = Tbl^T pb */
  double p0[3] = {0.0, 0.0, 0.0};
  if (b->numEleLoads > 0) {          /* computeReactions, ForceBeamColumn2d.cpp:407-462 */
    double wa = b->w[2] * b->loadFactor, wy = b->w[0] * b->loadFactor;
    p0[0] -= wa * b->L;
    double Vr = 0.5 * wy * b->L;
    p0[1] -= Vr; p0[2] -= Vr;
    if (b->has_point) {
      double P = b->pt[0] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
      double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
      p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
    }
    beam2_partial_p0(b, p0);
  }
  pl[0] += p0[0]; pl[1] += p0[1]; pl[4] += p0[2];      /* member loads in the local system, :493-497 */
  R[0] = cosTheta * pl[0] - sinTheta * pl[1];
  R[1] = sinTheta * pl[0] + cosTheta * pl[1];
  R[3] = cosTheta * pl[3] - sinTheta * pl[4];
  R[4] = sinTheta * pl[3] + cosTheta * pl[4];
  R[2] = pl[2]; R[5] = pl[5];
}
static void beam_form_end(const OrcBeam* b, double* K, double* R);
/* tangent and resisting force at the NODES: those of the element ends, pulled back through the rigid offsets
 * (K_node = To' K_end To, R_node = To' R_end with u_end = To u_node; the t02, t12, t35, t45 terms of
 * LinearCrdTransf2d::getGlobalStiffMatrix / getGlobalResistingForce, LinearCrdTransf2d.cpp) */
static void beam_form(const OrcBeam* b, double* K, double* R) {
  beam_form_end(b, K, R);
  if (!b->has_off) return;
  const double c[2][2] = {{-b->off[1], b->off[0]}, {-b->off[3], b->off[2]}};     /* To[3a + p][3a + 2] */
  if (K) {
    for (int i = 0; i < 6; i++) for (int a = 0; a < 2; a++) K[i * 6 + 3 * a + 2] += K[i * 6 + 3 * a] * c[a][0] + K[i * 6 + 3 * a + 1] * c[a][1];
    for (int j = 0; j < 6; j++) for (int a = 0; a < 2; a++) K[(3 * a + 2) * 6 + j] += c[a][0] * K[(3 * a) * 6 + j] + c[a][1] * K[(3 * a + 1) * 6 + j];
  }
  for (int a = 0; a < 2; a++) R[3 * a + 2] += c[a][0] * R[3 * a] + c[a][1] * R[3 * a + 1];
}
/* LinearCrdTransf2d::getGlobalStiffMatrix and getGlobalResistingForce at the element ends; K row-major 6x6 */
static void beam_form_end(const OrcBeam* b, double* K, double* R) {
  if (b->pdelta) { beam_form_pdelta(b, K, R); return; }
  if (b->corot) { beam_form_corot(b, K, R); return; }
  const double cosTheta = b->cosTheta, sinTheta = b->sinTheta, oneOverL = 1.0 / b->L;
  if (K) {
    const double* kb = b->kv;
    double kb00 = kb[0], kb10 = kb[1], kb20 = kb[2], kb01 = kb[3], kb11 = kb[4], kb21 = kb[5], kb02 = kb[6], kb12 = kb[7], kb22 = kb[8];
    double tmp[3][6], kg[6][6];
    double sl = sinTheta * oneOverL, cl = cosTheta * oneOverL;
    tmp[0][0] = -cosTheta * kb00 - sl * (kb01 + kb02); tmp[0][1] = -sinTheta * kb00 + cl * (kb01 + kb02);
    tmp[0][2] = kb01; tmp[0][3] = -tmp[0][0]; tmp[0][4] = -tmp[0][1]; tmp[0][5] = kb02;
    tmp[1][0] = -cosTheta * kb10 - sl * (kb11 + kb12); tmp[1][1] = -sinTheta * kb10 + cl * (kb11 + kb12);
    tmp[1][2] = kb11; tmp[1][3] = -tmp[1][0]; tmp[1][4] = -tmp[1][1]; tmp[1][5] = kb12;
    tmp[2][0] = -cosTheta * kb20 - sl * (kb21 + kb22); tmp[2][1] = -sinTheta * kb20 + cl * (kb21 + kb22);
    tmp[2][2] = kb21; tmp[2][3] = -tmp[2][0]; tmp[2][4] = -tmp[2][1]; tmp[2][5] = kb22;
    for (int c = 0; c < 6; c++) {
      kg[0][c] = -cosTheta * tmp[0][c] - sl * (tmp[1][c] + tmp[2][c]);
      kg[1][c] = -sinTheta * tmp[0][c] + cl * (tmp[1][c] + tmp[2][c]);
      kg[2][c] = tmp[1][c];
      kg[3][c] = -kg[0][c]; kg[4][c] = -kg[1][c];
      kg[5][c] = tmp[2][c];
    }
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) K[r * 6 + c] = kg[r][c];
  }
  double q0 = b->Se[0], q1 = b->Se[1], q2 = b->Se[2];
  double V = oneOverL * (q1 + q2);
  double pl[6] = { -q0, V, q1, q0, -V, q2 };
  double p0[3] = {0.0, 0.0, 0.0};
  if (b->numEleLoads > 0) {   /* computeReactions, ForceBeamColumn2d.cpp:407-425 */
    double wa = b->w[2] * b->loadFactor, wy = b->w[0] * b->loadFactor;
    p0[0] -= wa * b->L;
    double Vr = 0.5 * wy * b->L;
    p0[1] -= Vr; p0[2] -= Vr;
    if (b->has_point) {               /* Beam2dPointLoad, ForceBeamColumn2d.cpp:442-455 */
      double P = b->pt[0] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
      double V1 = P * (1.0 - aOverL), V2 = P * aOverL;
      p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
    }
    beam2_partial_p0(b, p0);
  }
  pl[0] += p0[0]; pl[1] += p0[1]; pl[4] += p0[2];
  R[0] = cosTheta * pl[0] - sinTheta * pl[1];
  R[1] = sinTheta * pl[0] + cosTheta * pl[1];
  R[3] = cosTheta * pl[3] - sinTheta * pl[4];
  R[4] = sinTheta * pl[3] + cosTheta * pl[4];
  R[2] = pl[2]; R[5] = pl[5];
}
/* ForceBeamColumn2d::commitState / revertToLastCommit, ForceBeamColumn2d.cpp:276-342 */
static void beam_commit(OrcBeam* b) {
  for (int i = 0; i < b->nip; i++) {
    memcpy(b->vscommit[i], b->vs[i], sizeof b->vs[i]);
    for (int f = 0; f < b->sec[i].nf; f++) uni_commit(&b->sec[i].mat[f]);
  }
  memcpy(b->kvcommit, b->kv, sizeof b->kv); memcpy(b->Secommit, b->Se, sizeof b->Se);
  memcpy(b->cubc, b->cub, sizeof b->cub);       /* CorotCrdTransf2d::commitState */
}
static void beam_revert(OrcBeam* b) {
  /* CorotCrdTransf2d::revertToLastCommit: ub = ubcommit, update() at the (already reverted) nodes' displacements */
  memcpy(b->cub, b->cubc, sizeof b->cub);
  for (int i = 0; i < b->nip; i++) {
    memcpy(b->vs[i], b->vscommit[i], sizeof b->vs[i]);
    sec_revert(&b->sec[i]);
    sec_set_trial(&b->sec[i], b->vs[i]);
    b->Ssr[i][0] = b->sec[i].s[0]; b->Ssr[i][1] = b->sec[i].s[1];
    sec_flex(&b->sec[i], b->fs[i]);
  }
  memcpy(b->Se, b->Secommit, sizeof b->Se); memcpy(b->kv, b->kvcommit, sizeof b->kv);
  b->initialFlag = 0;
}

/* ======================================================================== */
/* FiberSection3d (SRC/material/section/FiberSection3d.cpp), `section Fiber tag -GJ gj`: */
/* code = P, MZ, MY, T; elastic torsion (ElasticMaterial GJ)                   */
/* ======================================================================== */
typedef struct {
  int nf; double* y; double* z; double* A; OrcUni* mat; double yBar, zBar, GJ;
  double e[4], s[4], k[16];    /* kData column-major 4x4 */
} OrcSec3;

/* FiberSection3d::setTrialSectionDeformation, FiberSection3d.cpp:422-475 */
static int sec3_set_trial(OrcSec3* S, const double* d) {
  for (int i = 0; i < 4; i++) { S->e[i] = d[i]; S->s[i] = 0.0; }
  for (int i = 0; i < 16; i++) S->k[i] = 0.0;
  const double e0 = d[0], e1 = d[1], e2 = d[2], e3 = d[3];
  int res = 0;
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, z = S->z[i] - S->zBar, A = S->A[i];
    double strain = e0 - y * e1 + z * e2;
    res += uni_set_trial(&S->mat[i], strain);
    double tangent = S->mat[i].e, stress = S->mat[i].sig;
    double EA = tangent * A;
    S->k[0] += EA; S->k[1] += -y * EA; S->k[2] += z * EA;
    S->k[5] += y * y * EA; S->k[10] += z * z * EA; S->k[6] += -y * z * EA;
    double fs0 = stress * A;
    S->s[0] += fs0; S->s[1] += -y * fs0; S->s[2] += z * fs0;
  }
  S->k[4] = S->k[1]; S->k[8] = S->k[2]; S->k[9] = S->k[6];
  /* theTorsion->setTrial(e3, stress, tangent): ElasticMaterial */
  S->s[3] = S->GJ * e3; S->k[15] = S->GJ;
  return res;
}
/* FiberSection3d::revertToLastCommit, FiberSection3d.cpp:612-666 (sData[3] is left at 0) */
static void sec3_revert(OrcSec3* S) {
  for (int i = 0; i < 16; i++) S->k[i] = 0.0;
  for (int i = 0; i < 4; i++) S->s[i] = 0.0;
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, z = S->z[i] - S->zBar, A = S->A[i];
    uni_revert(&S->mat[i]);
    double value = S->mat[i].e * A, vas1 = -y * value, vas2 = z * value, vas1as2 = vas1 * z;
    S->k[0] += value; S->k[1] += vas1; S->k[2] += vas2;
    S->k[5] += vas1 * -y; S->k[6] += vas1as2; S->k[10] += vas2 * z;
    double fs0 = S->mat[i].sig * A;
    S->s[0] += fs0; S->s[1] += fs0 * -y; S->s[2] += fs0 * z;
  }
  S->k[4] = S->k[1]; S->k[8] = S->k[2]; S->k[9] = S->k[6];
  S->k[15] = S->GJ;
}
/* SectionForceDeformation::getSectionFlexibility -> Matrix::Invert -> cmx_inv4 (matrix/routines/invGL4.c,
 * a cofactor expansion).  ks is block diagonal here (P-Mz-My block, torsion): the cofactor
 * expansion of the 4x4 reduces to that of the 3x3 block times k33 over det3*k33, so the block is
 * inverted with the 3x3 cofactor formula and the torsion entry by division (same to rounding). */
static void sec3_flex(const double* k, double* f) {
  double a[9], ai[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) a[r + 3 * c] = k[r + 4 * c];
  inv3(a, ai);
  for (int i = 0; i < 16; i++) f[i] = 0.0;
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) f[r + 4 * c] = ai[r + 3 * c];
  f[15] = 1.0 / k[15];
}
static void sec3_initial_flex(const OrcSec3* S, double* f) {   /* FiberSection3d::getInitialTangent + Invert */
  double k[16]; for (int i = 0; i < 16; i++) k[i] = 0.0;
  for (int i = 0; i < S->nf; i++) {
    const double y = S->y[i] - S->yBar, z = S->z[i] - S->zBar, A = S->A[i];
    const double EA = uni_initial_tangent(&S->mat[i]) * A;
    const double vas2 = z * EA;
    k[0] += EA; k[1] += -y * EA; k[2] += z * EA; k[5] += y * y * EA; k[6] += -y * z * EA; k[10] += vas2 * z;
  }
  k[4] = k[1]; k[8] = k[2]; k[9] = k[6]; k[15] = S->GJ;
  sec3_flex(k, f);
}

/* ======================================================================== */
/* ForceBeamColumn3d (SRC/element/Frame/Other/Force/ForceBeamColumn3d.cpp)    */
/* with LinearCrdTransf3d (no offsets) and LobattoBeamIntegration              */
/* ======================================================================== */
typedef struct {
  int nip, maxIters; double tol;
  OrcSec3 sec[ORC_MAXSEC];
  double L, R[3][3];
  int initialFlag;
  double kv[36], Se[6], kvcommit[36], Secommit[6];       /* kv column-major 6x6 */
  double fs[ORC_MAXSEC][16], vs[ORC_MAXSEC][4], Ssr[ORC_MAXSEC][4], vscommit[ORC_MAXSEC][4];
  /* `eleLoad -beamUniform wy wz wa`: see OrcBeam */
  int has_load, numEleLoads; double w[3], loadFactor;
  /* `eleLoad -beamPoint Py Pz xL [N]` (Beam3dPointLoad): pt = {Py, Pz, N, aOverL} */
  int has_point; double pt[4];
  /* Beam3dPartialUniformLoad: pq = wy_a, wy_b, wAxial_a, wAxial_b, aOverL, bOverL, wz_a, wz_b */
  int has_partial; double pq[8];
  int user_rule; double rxi[ORC_MAXSEC], rwt[ORC_MAXSEC];     /* see OrcBeam */
  /* geomTransf PDelta (PDeltaCrdTransf3d.cpp:200-249): ul17, ul28 as of the element's last update() -- ForceBeamColumn3d's
   * getTangentStiff / getResistingForce do NOT refresh them (ForceBeamColumn3d.cpp:404,555) */
  int pdelta; double ul17, ul28;
  int has_off; double off[6];      /* -jntOffset dXi dYi dZi dXj dYj dZj (LinearCrdTransf3d.cpp / PDeltaCrdTransf3d.cpp) */
} OrcBeam3;

/* LinearCrdTransf3d::initialize -> computeElemtLengthAndOrient + getLocalAxes, LinearCrdTransf3d.cpp:203-330 */
static int crd3d_init(OrcBeam3* b, const double* xi, const double* xj, const double* vecxz) {
  double dx[3] = { xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2] };
  b->L = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
  if (b->L == 0.0) return -2;
  for (int i = 0; i < 3; i++) b->R[0][i] = dx[i] / b->L;
  const double* v = vecxz; const double* x = b->R[0];
  double y[3] = { v[1] * x[2] - v[2] * x[1], v[2] * x[0] - v[0] * x[2], v[0] * x[1] - v[1] * x[0] };
  double ynorm = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
  if (ynorm == 0) return -3;
  for (int i = 0; i < 3; i++) y[i] /= ynorm;
  double z[3] = { x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0] };
  for (int i = 0; i < 3; i++) { b->R[1][i] = y[i]; b->R[2][i] = z[i]; }
  return 0;
}
/* LinearCrdTransf3d::getBasicTrialDisp / getBasicIncrDeltaDisp, LinearCrdTransf3d.cpp:344-421 */
static void crd3d_basic(const OrcBeam3* b, const double* ug, double* ub) {
  double ul[12];
  for (int blk = 0; blk < 4; blk++)
    for (int r = 0; r < 3; r++)
      ul[3 * blk + r] = b->R[r][0] * ug[3 * blk] + b->R[r][1] * ug[3 * blk + 1] + b->R[r][2] * ug[3 * blk + 2];
  double oneOverL = 1.0 / b->L, tmp;
  ub[0] = ul[6] - ul[0];
  tmp = oneOverL * (ul[1] - ul[7]);
  ub[1] = ul[5] + tmp; ub[2] = ul[11] + tmp;
  tmp = oneOverL * (ul[8] - ul[2]);
  ub[3] = ul[4] + tmp; ub[4] = ul[10] + tmp;
  ub[5] = ul[9] - ul[3];
}
/* Matrix::Invert of the 6x6 element flexibility (cmx_inv6, matrix/routines/invGL6.c: a generated
 * cofactor expansion).  The torsion row/column is uncoupled here; the 5x5 block is inverted by
 * Gauss-Jordan elimination with partial pivoting -- agreement with the cofactor expansion is to
 * rounding times the block's condition number. f, kv column-major 6x6. */
static int inv6_flex(const double* f, double* kv) {
  double a[5][10];
  for (int r = 0; r < 5; r++) { for (int c = 0; c < 5; c++) { a[r][c] = f[r + 6 * c]; a[r][5 + c] = (r == c) ? 1.0 : 0.0; } }
  for (int c = 0; c < 5; c++) {
    int p = c; double big = fabs(a[c][c]);
    for (int r = c + 1; r < 5; r++) if (fabs(a[r][c]) > big) { big = fabs(a[r][c]); p = r; }
    if (big == 0.0) return -1;
    if (p != c) for (int q = 0; q < 10; q++) { double t = a[c][q]; a[c][q] = a[p][q]; a[p][q] = t; }
    double piv = 1.0 / a[c][c];
    for (int q = 0; q < 10; q++) a[c][q] *= piv;
    for (int r = 0; r < 5; r++) if (r != c) { double m = a[r][c]; if (m != 0.0) for (int q = 0; q < 10; q++) a[r][q] -= m * a[c][q]; }
  }
  for (int i = 0; i < 36; i++) kv[i] = 0.0;
  for (int r = 0; r < 5; r++) for (int c = 0; c < 5; c++) kv[r + 6 * c] = a[r][5 + c];
  kv[35] = 1.0 / f[35];
  return 0;
}
static double norm6(const double* v) { double s = 0.0; for (int i = 0; i < 6; i++) s += v[i] * v[i]; return sqrt(s); }

/* ForceBeamColumn3d::update, ForceBeamColumn3d.cpp:587-1056 (no element loads; isTorsion = true) */
/* displacements of the element ends from those of the nodes (joint offsets): u_end = u + theta x offset */
static void beam3_end_disp(const OrcBeam3* b, double* ug) {
  if (!b->has_off) return;
  for (int a = 0; a < 2; a++) {
    const double* d = b->off + 3 * a; double* u = ug + 6 * a;
    const double tx = u[3], ty = u[4], tz = u[5];
    u[0] += ty * d[2] - tz * d[1]; u[1] += tz * d[0] - tx * d[2]; u[2] += tx * d[1] - ty * d[0];
  }
}
static int beam3_update(OrcBeam3* b, const double* ug_, const double* dug_) {
  double ug[12], dug[12];
  memcpy(ug, ug_, sizeof ug); memcpy(dug, dug_, sizeof dug);
  beam3_end_disp(b, ug); beam3_end_disp(b, dug);
  if (b->pdelta) {   /* crdTransf->update(), PDeltaCrdTransf3d.cpp:200-249 (no offsets) */
    const double ul1 = b->R[1][0] * ug[0] + b->R[1][1] * ug[1] + b->R[1][2] * ug[2];
    const double ul2 = b->R[2][0] * ug[0] + b->R[2][1] * ug[1] + b->R[2][2] * ug[2];
    const double ul7 = b->R[1][0] * ug[6] + b->R[1][1] * ug[7] + b->R[1][2] * ug[8];
    const double ul8 = b->R[2][0] * ug[6] + b->R[2][1] * ug[7] + b->R[2][2] * ug[8];
    b->ul17 = ul1 - ul7; b->ul28 = ul2 - ul8;
  }
  double v[6], dv[6], vin[6];
  crd3d_basic(b, ug, v);
  crd3d_basic(b, dug, dv);
  if (b->initialFlag != 0 && norm6(dv) <= DBL_EPSILON && b->numEleLoads == 0) return 0;
  for (int i = 0; i < 6; i++) vin[i] = v[i] - dv[i];
  const double L = b->L;
  double xi[ORC_MAXSEC], wt[ORC_MAXSEC];
  if (b->user_rule) { memcpy(xi, b->rxi, sizeof(double) * b->nip); memcpy(wt, b->rwt, sizeof(double) * b->nip); } else lobatto(b->nip, xi, wt);
  double vr[6], f[36], dSe[6], SeTrial[6], kvTrial[36], dvTrial[6], dvToDo[6];
  double vsSub[ORC_MAXSEC][4], fsSub[ORC_MAXSEC][16], SsrSub[ORC_MAXSEC][4];
  int numSubdivide = 1, converged = 0;
  for (int i = 0; i < 6; i++) { dvToDo[i] = dv[i]; dvTrial[i] = dvToDo[i]; }
  const double factor = 10.0;
  const int maxSubdivisions = 10;
  while (!converged && numSubdivide <= maxSubdivisions) {
    for (int l = 0; l < 3; l++) {
      memcpy(SeTrial, b->Se, sizeof SeTrial); memcpy(kvTrial, b->kv, sizeof kvTrial);
      for (int i = 0; i < b->nip; i++) {
        memcpy(vsSub[i], b->vs[i], sizeof vsSub[i]); memcpy(fsSub[i], b->fs[i], sizeof fsSub[i]);
        memcpy(SsrSub[i], b->Ssr[i], sizeof SsrSub[i]);
      }
      for (int i = 0; i < 6; i++) dSe[i] = 0.0;
      for (int j = 0; j < 6; j++) for (int i = 0; i < 6; i++) dSe[i] += kvTrial[i + 6 * j] * dvTrial[j];
      for (int i = 0; i < 6; i++) SeTrial[i] += dSe[i];
      int numIters = b->maxIters;
      if (l == 1) numIters = 10 * b->maxIters;
      for (int j = 0; j < numIters; j++) {
        for (int i = 0; i < 36; i++) f[i] = 0.0;
        for (int i = 0; i < 6; i++) vr[i] = 0.0;
        for (int i = 0; i < b->nip; i++) {
          OrcSec3* S = &b->sec[i];
          double Ss[4], dSs[4], dvs[4], fb[24];   /* fb (4 x 6) column-major */
          double xL = xi[i], xL1 = xL - 1.0, wtL = wt[i] * L;
          Ss[0] = SeTrial[0];
          Ss[1] = xL1 * SeTrial[1] + xL * SeTrial[2];
          Ss[2] = xL1 * SeTrial[3] + xL * SeTrial[4];
          Ss[3] = SeTrial[5];
          if (b->numEleLoads > 0) {   /* computeSectionForces, ForceBeamColumn3d.cpp:1197-1215 (Beam3dUniformLoad) */
            double x = xi[i] * L;
            double wy = b->w[0] * b->loadFactor, wz = b->w[1] * b->loadFactor, wa = b->w[2] * b->loadFactor;
            Ss[0] += wa * (L - x);
            Ss[1] += wy * 0.5 * x * (x - L);
            Ss[2] += wz * 0.5 * x * (L - x);
            if (b->has_point) {       /* Beam3dPointLoad, ForceBeamColumn3d.cpp:1314-1373 */
              double Py = b->pt[0] * b->loadFactor, Pz = b->pt[1] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
              double a = aOverL * L;
              double Vy1 = Py * (1.0 - aOverL), Vy2 = Py * aOverL, Vz1 = Pz * (1.0 - aOverL), Vz2 = Pz * aOverL;
              if (x <= a) { Ss[0] += N; Ss[1] -= x * Vy1; Ss[2] += x * Vz1; }
              else { Ss[1] -= (L - x) * Vy2; Ss[2] += (L - x) * Vz2; }
            }
            if (b->has_partial) {     /* Beam3dPartialUniformLoad, ForceBeamColumn3d.cpp:1224-1313 */
              const double lf = b->loadFactor;
              double wy = b->pq[0] * lf, wz = b->pq[6] * lf, wa = b->pq[2] * lf, a = b->pq[4] * L, bb = b->pq[5] * L;
              double wyb = b->pq[1] * lf, wzb = b->pq[7] * lf, wab = b->pq[3] * lf;
              double Fa = wa * (bb - a) + 0.5 * (wab - wa) * (bb - a);
              double Fy = wy * (bb - a), Fz = wz * (bb - a);
              double c = a + 0.5 * (bb - a);
              double VyI = Fy * (1 - c / L), VyJ = Fy * c / L, VzI = Fz * (1 - c / L), VzJ = Fz * c / L;
              Fy = 0.5 * (wyb - wy) * (bb - a); Fz = 0.5 * (wzb - wz) * (bb - a);
              c = a + 2.0 / 3.0 * (bb - a);
              VyI += Fy * (1 - c / L); VyJ += Fy * c / L; VzI += Fz * (1 - c / L); VzJ += Fz * c / L;
              if (x <= a) { Ss[0] += Fa; Ss[1] -= VyI * x; Ss[2] += VzI * x; }
              else if (x >= bb) { Ss[1] += VyJ * (x - L); Ss[2] -= VzJ * (x - L); }
              else {
                double wyy = wy + (wyb - wy) / (bb - a) * (x - a), wzz = wz + (wzb - wz) / (bb - a) * (x - a);
                Ss[0] += Fa - wa * (x - a) - 0.5 * (wab - wa) / (bb - a) * (x - a) * (x - a);
                Ss[1] += -VyI * x + 0.5 * wy * (x - a) * (x - a) + 0.5 * (wyy - wy) * (x - a) * (x - a) / 3.0;
                Ss[2] += VzI * x - 0.5 * wz * (x - a) * (x - a) - 0.5 * (wzz - wz) * (x - a) * (x - a) / 3.0;
              }
            }
          }
          for (int q = 0; q < 4; q++) dSs[q] = Ss[q] - SsrSub[i][q];
          const double* fuse;
          double fs0[16];
          if (l == 0) fuse = fsSub[i];
          else if (l == 2) { if (j == 0) { sec3_initial_flex(S, fs0); fuse = fs0; } else fuse = fsSub[i]; }
          else { sec3_initial_flex(S, fs0); fuse = fs0; }
          for (int q = 0; q < 4; q++) dvs[q] = 0.0;
          for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) dvs[r] += fuse[r + 4 * c] * dSs[c];
          if (b->initialFlag != 0) for (int q = 0; q < 4; q++) vsSub[i][q] += dvs[q];
          if (sec3_set_trial(S, vsSub[i]) < 0) return -1;
          for (int q = 0; q < 4; q++) SsrSub[i][q] = S->s[q];
          sec3_flex(S->k, fsSub[i]);
          for (int q = 0; q < 4; q++) dSs[q] = Ss[q] - SsrSub[i][q];
          for (int q = 0; q < 4; q++) dvs[q] = 0.0;
          for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) dvs[r] += fsSub[i][r + 4 * c] * dSs[c];
          /* fb = fs * b * wtL ; code = {P, MZ, MY, T} */
          for (int q = 0; q < 24; q++) fb[q] = 0.0;
          const double* fSec = fsSub[i];
          for (int jj = 0; jj < 4; jj++) fb[jj + 4 * 0] += fSec[jj + 4 * 0] * wtL;
          for (int jj = 0; jj < 4; jj++) { double tmp = fSec[jj + 4 * 1] * wtL; fb[jj + 4 * 1] += xL1 * tmp; fb[jj + 4 * 2] += xL * tmp; }
          for (int jj = 0; jj < 4; jj++) { double tmp = fSec[jj + 4 * 2] * wtL; fb[jj + 4 * 3] += xL1 * tmp; fb[jj + 4 * 4] += xL * tmp; }
          for (int jj = 0; jj < 4; jj++) fb[jj + 4 * 5] += fSec[jj + 4 * 3] * wtL;
          /* f += b^T fb */
          for (int jj = 0; jj < 6; jj++) f[0 + 6 * jj] += fb[0 + 4 * jj];
          for (int jj = 0; jj < 6; jj++) { double tmp = fb[1 + 4 * jj]; f[1 + 6 * jj] += xL1 * tmp; f[2 + 6 * jj] += xL * tmp; }
          for (int jj = 0; jj < 6; jj++) { double tmp = fb[2 + 4 * jj]; f[3 + 6 * jj] += xL1 * tmp; f[4 + 6 * jj] += xL * tmp; }
          for (int jj = 0; jj < 6; jj++) f[5 + 6 * jj] += fb[3 + 4 * jj];
          /* vr += b^T (vs + dvs) wtL */
          for (int q = 0; q < 4; q++) dvs[q] += vsSub[i][q];
          { double dei = dvs[0] * wtL; vr[0] += dei; }
          { double dei = dvs[1] * wtL; vr[1] += xL1 * dei; vr[2] += xL * dei; }
          { double dei = dvs[2] * wtL; vr[3] += xL1 * dei; vr[4] += xL * dei; }
          { double dei = dvs[3] * wtL; vr[5] += dei; }
        }
        if (inv6_flex(f, kvTrial) < 0) return -1;
        for (int i = 0; i < 6; i++) { dv[i] = vin[i]; dv[i] += dvTrial[i]; dv[i] -= vr[i]; }
        for (int i = 0; i < 6; i++) dSe[i] = 0.0;
        for (int c = 0; c < 6; c++) for (int r = 0; r < 6; r++) dSe[r] += kvTrial[r + 6 * c] * dv[c];
        double dW = 0.0;
        for (int i = 0; i < 6; i++) dW += dv[i] * dSe[i];
        for (int i = 0; i < 6; i++) SeTrial[i] += dSe[i];
        if (fabs(dW) < b->tol) {
          for (int i = 0; i < 6; i++) { dvToDo[i] -= dvTrial[i]; vin[i] += dvTrial[i]; }
          if (norm6(dvToDo) <= DBL_EPSILON) converged = 1;
          else { for (int i = 0; i < 6; i++) dvTrial[i] = dvToDo[i]; numSubdivide = 1; }
          memcpy(b->kv, kvTrial, sizeof kvTrial); memcpy(b->Se, SeTrial, sizeof SeTrial);
          for (int k = 0; k < b->nip; k++) {
            memcpy(b->vs[k], vsSub[k], sizeof vsSub[k]); memcpy(b->fs[k], fsSub[k], sizeof fsSub[k]);
            memcpy(b->Ssr[k], SsrSub[k], sizeof SsrSub[k]);
          }
          j = numIters + 1; l = 4;
        } else {
          if (j == (numIters - 1) && (l == 2)) { for (int i = 0; i < 6; i++) dvTrial[i] /= factor; numSubdivide++; }
        }
      }
    }
  }
  if (!converged) return -1;
  b->initialFlag = 1;
  return 0;
}

/* LinearCrdTransf3d::getGlobalStiffMatrix (no offsets, :767-926) and getGlobalResistingForce (:699-765);
 * K row-major 12x12 */
static void beam3_form_end(const OrcBeam3* b, double* K, double* Rg);
/* tangent and resisting force at the NODES: K_node = To' K_end To, R_node = To' R_end with u_end = To u_node (the joint-offset
 * terms of LinearCrdTransf3d::getGlobalStiffMatrix / getGlobalResistingForce) */
static void beam3_form(const OrcBeam3* b, double* K, double* Rg) {
  beam3_form_end(b, K, Rg);
  if (!b->has_off) return;
  for (int a = 0; a < 2; a++) {
    const double* d = b->off + 3 * a;
    const double C[3][3] = {{0.0, d[2], -d[1]}, {-d[2], 0.0, d[0]}, {d[1], -d[0], 0.0}};      /* u_end = u + C theta */
    if (K) {
      for (int i = 0; i < 12; i++) for (int j = 0; j < 3; j++)
        K[i * 12 + 6 * a + 3 + j] += K[i * 12 + 6 * a] * C[0][j] + K[i * 12 + 6 * a + 1] * C[1][j] + K[i * 12 + 6 * a + 2] * C[2][j];
      for (int i = 0; i < 12; i++) for (int j = 0; j < 3; j++)
        K[(6 * a + 3 + j) * 12 + i] += C[0][j] * K[(6 * a) * 12 + i] + C[1][j] * K[(6 * a + 1) * 12 + i] + C[2][j] * K[(6 * a + 2) * 12 + i];
    }
    for (int j = 0; j < 3; j++) Rg[6 * a + 3 + j] += C[0][j] * Rg[6 * a] + C[1][j] * Rg[6 * a + 1] + C[2][j] * Rg[6 * a + 2];
  }
}
static void beam3_form_end(const OrcBeam3* b, double* K, double* Rg) {
  const double oneOverL = 1.0 / b->L;
  const double (*R)[3] = b->R;
  if (K) {
    double kb[6][6], kl[12][12], tmp[12][12];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) kb[i][j] = b->kv[i + 6 * j];
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
    if (b->pdelta) {   /* PDeltaCrdTransf3d::getGlobalStiffMatrix, PDeltaCrdTransf3d.cpp:873-881 */
      const double NoverL = b->Se[0] * oneOverL;
      kl[1][1] += NoverL; kl[2][2] += NoverL; kl[7][7] += NoverL; kl[8][8] += NoverL;
      kl[1][7] -= NoverL; kl[7][1] -= NoverL; kl[2][8] -= NoverL; kl[8][2] -= NoverL;
    }
    for (int m = 0; m < 12; m++)
      for (int blk = 0; blk < 4; blk++)
        for (int c = 0; c < 3; c++)
          tmp[m][3 * blk + c] = kl[m][3 * blk] * R[0][c] + kl[m][3 * blk + 1] * R[1][c] + kl[m][3 * blk + 2] * R[2][c];
    for (int m = 0; m < 12; m++)
      for (int blk = 0; blk < 4; blk++)
        for (int c = 0; c < 3; c++)
          K[(3 * blk + c) * 12 + m] = R[0][c] * tmp[3 * blk][m] + R[1][c] * tmp[3 * blk + 1][m] + R[2][c] * tmp[3 * blk + 2][m];
  }
  const double* q = b->Se;
  double pl[12];
  pl[0] = -q[0]; pl[1] = oneOverL * (q[1] + q[2]); pl[2] = -oneOverL * (q[3] + q[4]); pl[3] = -q[5];
  pl[4] = q[3]; pl[5] = q[1]; pl[6] = q[0]; pl[7] = -pl[1]; pl[8] = -pl[2]; pl[9] = q[5]; pl[10] = q[4]; pl[11] = q[2];
  double p0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (b->numEleLoads > 0) {   /* computeReactions, ForceBeamColumn3d.cpp:419-431 */
    double wy = b->w[0] * b->loadFactor, wz = b->w[1] * b->loadFactor, wa = b->w[2] * b->loadFactor;
    p0[0] -= wa * b->L;
    double Vr = 0.5 * wy * b->L;
    p0[1] -= Vr; p0[2] -= Vr;
    Vr = 0.5 * wz * b->L;
    p0[3] -= Vr; p0[4] -= Vr;
    if (b->has_point) {               /* Beam3dPointLoad, ForceBeamColumn3d.cpp:457-475 */
      double Py = b->pt[0] * b->loadFactor, Pz = b->pt[1] * b->loadFactor, N = b->pt[2] * b->loadFactor, aOverL = b->pt[3];
      double V1 = Py * (1.0 - aOverL), V2 = Py * aOverL;
      p0[0] -= N; p0[1] -= V1; p0[2] -= V2;
      V1 = Pz * (1.0 - aOverL); V2 = Pz * aOverL;
      p0[3] -= V1; p0[4] -= V2;
    }
    if (b->has_partial) {             /* Beam3dPartialUniformLoad, ForceBeamColumn3d.cpp:432-456 */
      const double lf = b->loadFactor, L = b->L;
      double wy = b->pq[0] * lf, wz = b->pq[6] * lf, wa = b->pq[2] * lf, a = b->pq[4] * L, bb = b->pq[5] * L;
      double wyb = b->pq[1] * lf, wzb = b->pq[7] * lf, wab = b->pq[3] * lf;
      p0[0] -= wa * (bb - a) + 0.5 * (wab - wa) * (bb - a);
      double c = a + 0.5 * (bb - a);
      double Fy = wy * (bb - a);
      p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
      double Fz = wz * (bb - a);
      p0[3] -= Fz * (1 - c / L); p0[4] -= Fz * c / L;
      c = a + 2.0 / 3.0 * (bb - a);
      Fy = 0.5 * (wyb - wy) * (bb - a);
      p0[1] -= Fy * (1 - c / L); p0[2] -= Fy * c / L;
      Fz = 0.5 * (wzb - wz) * (bb - a);
      p0[3] -= Fz * (1 - c / L); p0[4] -= Fz * c / L;
    }
  }
  pl[0] += p0[0]; pl[1] += p0[1]; pl[7] += p0[2]; pl[2] += p0[3]; pl[8] += p0[4];   /* LinearCrdTransf3d.cpp:727-731 */
  if (b->pdelta) {   /* PDeltaCrdTransf3d::getGlobalResistingForce, PDeltaCrdTransf3d.cpp:784-790 */
    double NoverL = b->ul17 * q[0] * oneOverL;
    pl[1] += NoverL; pl[7] -= NoverL;
    NoverL = b->ul28 * q[0] * oneOverL;
    pl[2] += NoverL; pl[8] -= NoverL;
  }
  for (int blk = 0; blk < 4; blk++)
    for (int c = 0; c < 3; c++)
      Rg[3 * blk + c] = R[0][c] * pl[3 * blk] + R[1][c] * pl[3 * blk + 1] + R[2][c] * pl[3 * blk + 2];
}
/* ForceBeamColumn3d::commitState / revertToLastCommit, ForceBeamColumn3d.cpp:279-345 */
static void beam3_commit(OrcBeam3* b) {
  for (int i = 0; i < b->nip; i++) {
    memcpy(b->vscommit[i], b->vs[i], sizeof b->vs[i]);
    for (int f = 0; f < b->sec[i].nf; f++) uni_commit(&b->sec[i].mat[f]);
  }
  memcpy(b->kvcommit, b->kv, sizeof b->kv); memcpy(b->Secommit, b->Se, sizeof b->Se);
}
static void beam3_revert(OrcBeam3* b) {
  for (int i = 0; i < b->nip; i++) {
    memcpy(b->vs[i], b->vscommit[i], sizeof b->vs[i]);
    sec3_revert(&b->sec[i]);
    sec3_set_trial(&b->sec[i], b->vs[i]);
    for (int q = 0; q < 4; q++) b->Ssr[i][q] = b->sec[i].s[q];
    sec3_flex(b->sec[i].k, b->fs[i]);
  }
  memcpy(b->Se, b->Secommit, sizeof b->Se); memcpy(b->kv, b->kvcommit, sizeof b->kv);
  b->initialFlag = 0;
}

/* ======================================================================== */
/* the model: Domain + AnalysisModel + LinearSOE flattened                    */
/* ======================================================================== */
typedef struct {
  int kind;          /* ORC_ELE_* */
  int tag;
  int nen, ndf_e;    /* nodes, dofs per node used by the element */
  int node[8];       /* node indices (into model arrays, tag-sorted) */
  int mat;           /* material index */
  double par[16];    /* brick: b1,b2,b3 ; quad: thickness,type,pressure,rho,b1,b2 ; beams: see orc_add_element */
  OrcGP gp[8];
  int nip;
  OrcBeam* beam;     /* ORC_ELE_FBC2D */
  OrcBeam3* beam3;   /* ORC_ELE_FBC3D */
  double* Kc;        /* Element::Kc (committed tangent), row-major nd x nd; allocated while betaKc != 0 */
} OrcEle;

typedef struct { int tag, nf; double* y; double* z; double* A; int* mat; double GJ; int agg; } OrcSecDef;

typedef struct {
  int ndm, ndf;
  int nn; int* node_tag; double* crd; /* [nn][ndm], ascending tag (MapOfTaggedObjects order, Domain.cpp:98) */
  double* trial; double* commit_disp; /* [nn][ndf] */
  double* incr;                       /* [nn][ndf] Node::getIncrDeltaDisp (Node.cpp:435) */
  double* mass;                       /* [nn][ndf] diagonal of Node::mass (`mass` command) */
  double* vel; double* acc; double* velc; double* accc;   /* trial / committed velocity, acceleration */
  double alphaM;                      /* Node::setRayleighDampingFactor */
  double e_alphaM, betaK, betaK0, betaKc;   /* Element::setRayleighDampingFactors (Domain::setRayleighDampingFactors) */
  double c1, c2, c3;                  /* TransientIntegrator coefficients (Newmark.cpp:117-140); 1,0,0 = static */
  int nuni; int* uni_tag; int* uni_kind; double* uni_par;   /* uniaxial materials [nuni][12] */
  int nsec; OrcSecDef* sec;           /* fibre section definitions */
  double* load;                       /* [nn][ndf] reference nodal loads (the current pattern, Linear series) */
  double* cload;                      /* [nn][ndf] loads of the patterns frozen by loadConst (NULL: none) */
  int* fixed;                         /* [nn][ndf] 1 when an SP_Constraint holds the dof */
  int* node_ndf;                      /* [nn] dofs the node really has (<= ndf): a FourNodeQuad's 2-dof nodes next to 3-dof frame nodes */
  int nmp; int* mp;                   /* equalDOF: (retained node, constrained node, dof) index triples */
  int nmat; int* mat_tag; int* mat_kind; double* mat_par; /* [nmat][8] */
  int ne, ecap; OrcEle* ele;          /* ascending element tag after setup */
  /* analysis side */
  int neq; int* id;                   /* [nn][ndf] equation numbers (DOF_Group::myID) */
  int soe_kind;                       /* 0 SparseGenCol (CSC), 1 SparseGenRow (CSR) */
  int* ptr; int* idx; int nnz;        /* colStartA/rowA or rowStartA/colA */
  double* A; double* B;
  double lambda, lambda_c;   /* load factor (Domain::currentTime under LoadControl) and its committed value */
  int ele_loads_const; double ele_lambda;   /* loadConst: the element loads keep the factor they had (LoadPattern::applyLoad, isConstant) */
  /* `system BandGeneral` (2) / `system ProfileSPD` (3) on top of the column graph: orc_set_store */
  int store_kind, numSubD, numSuperD, profileSize; int* iDiagLoc; double* Astore; long long astore_size;
} OrcModel;

static int find_node(const OrcModel* m, int tag) {
  int lo = 0, hi = m->nn - 1;
  while (lo <= hi) { int mid = (lo + hi) / 2; if (m->node_tag[mid] == tag) return mid; if (m->node_tag[mid] < tag) lo = mid + 1; else hi = mid - 1; }
  return -1;
}

/* nodes must be given with ascending tags (the Domain iterates its std::map that way) */
void* orc_model_new(int ndm, int ndf, int nn, const int* tags, const double* crd) {
  OrcModel* m = (OrcModel*)calloc(1, sizeof *m);
  m->ndm = ndm; m->ndf = ndf; m->nn = nn;
  m->node_tag = (int*)malloc(sizeof(int) * nn); memcpy(m->node_tag, tags, sizeof(int) * nn);
  for (int i = 1; i < nn; i++) if (tags[i] <= tags[i - 1]) { free(m->node_tag); free(m); return NULL; }
  m->crd = (double*)malloc(sizeof(double) * nn * ndm); memcpy(m->crd, crd, sizeof(double) * nn * ndm);
  m->trial = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->commit_disp = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->incr = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->mass = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->vel = (double*)calloc((size_t)nn * ndf, sizeof(double)); m->acc = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->velc = (double*)calloc((size_t)nn * ndf, sizeof(double)); m->accc = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->c1 = 1.0; m->c2 = 0.0; m->c3 = 0.0;
  m->load = (double*)calloc((size_t)nn * ndf, sizeof(double));
  m->fixed = (int*)calloc((size_t)nn * ndf, sizeof(int));
  m->node_ndf = (int*)malloc(sizeof(int) * (nn > 0 ? nn : 1));
  for (int i = 0; i < nn; i++) m->node_ndf[i] = ndf;
  m->mat_tag = NULL; m->nmat = 0;
  return m;
}
/* a node created under another `model -ndf`: it has nd <= ndf dofs; its DOF_Group::myID has nd entries (DOF_Group.cpp),
 * the missing ones never get an equation.  The nodal arrays keep the model's stride. */
int orc_set_node_ndf(void* h, int nodeTag, int nd) {
  OrcModel* m = (OrcModel*)h; int n = find_node(m, nodeTag);
  if (n < 0 || nd < 1 || nd > m->ndf) return -1;
  m->node_ndf[n] = nd; return 0;
}
int orc_fix(void* h, int nodeTag, int dof) {
  OrcModel* m = (OrcModel*)h; int n = find_node(m, nodeTag);
  if (n < 0 || dof < 0 || dof >= m->ndf) return -1;
  m->fixed[n * m->ndf + dof] = 1; return 0;
}
/* `equalDOF rNode cNode dofs...`: an MP_Constraint with an identity constraint matrix (same dofs on both nodes);
 * PlainHandler gives the constrained dofs id -4 (PlainHandler.cpp:129-176), the numberer then hands them the
 * retained dof's equation (PlainNumberer.cpp:111-142, DOF_Numberer.cpp:151-190) */
int orc_equal_dof(void* h, int rTag, int cTag, int n, const int* dofs) {
  OrcModel* m = (OrcModel*)h;
  int r = find_node(m, rTag), c = find_node(m, cTag);
  if (r < 0 || c < 0) return -1;
  m->mp = (int*)realloc(m->mp, sizeof(int) * 3 * (m->nmp + n));
  for (int i = 0; i < n; i++) { m->mp[3 * m->nmp] = r; m->mp[3 * m->nmp + 1] = c; m->mp[3 * m->nmp + 2] = dofs[i]; m->nmp++; }
  return 0;
}

int orc_add_nd_material(void* h, int tag, int kind, const double* p) {
  OrcModel* m = (OrcModel*)h;
  m->mat_tag = (int*)realloc(m->mat_tag, sizeof(int) * (m->nmat + 1));
  m->mat_kind = (int*)realloc(m->mat_kind, sizeof(int) * (m->nmat + 1));
  m->mat_par = (double*)realloc(m->mat_par, sizeof(double) * 8 * (m->nmat + 1));
  m->mat_tag[m->nmat] = tag; m->mat_kind[m->nmat] = kind;
  memset(m->mat_par + 8 * m->nmat, 0, 8 * sizeof(double));
  memcpy(m->mat_par + 8 * m->nmat, p, sizeof(double) * (kind == ORC_MAT_J2 ? 8 : 3));   /* J2: ..., eta, rho */
  m->nmat++; return 0;
}
/* ---- transient analysis (Newmark, displacement form; nodal masses, mass-proportional damping) ---- */
int orc_set_mass(void* h, int nodeTag, const double* mv) {
  OrcModel* m = (OrcModel*)h; int n = find_node(m, nodeTag); if (n < 0) return -1;
  for (int i = 0; i < m->ndf; i++) m->mass[n * m->ndf + i] = mv[i];
  return 0;
}
void orc_set_alphaM(void* h, double a) { ((OrcModel*)h)->alphaM = a; }
void orc_set_transient(void* h, double c1, double c2, double c3) { OrcModel* m = (OrcModel*)h; m->c1 = c1; m->c2 = c2; m->c3 = c3; }
/* Newmark::newStep, displacement unknown (Newmark.cpp:150-160): Udot = a1*Udot + a2*Udotdot(old),
 * Udotdot = a4*Udotdot + a3*Udot(old); Vector::addVector(thisFact, other, otherFact) */
void orc_newmark_predict(void* h, double a1, double a2, double a3, double a4) {
  OrcModel* m = (OrcModel*)h;
  for (int i = 0; i < m->nn * m->ndf; i++) {
    if (m->id[i] < 0) continue;
    double v0 = m->vel[i], ac0 = m->acc[i];
    m->vel[i] = v0 * a1 + ac0 * a2;
    m->acc[i] = ac0 * a4 + v0 * a3;
  }
}
/* AnalysisModel::setVel / setAccel: trial velocities and accelerations, [nn][ndf] */
void orc_set_vel_accel(void* h, const double* v, const double* a) {
  OrcModel* m = (OrcModel*)h;
  memcpy(m->vel, v, sizeof(double) * m->nn * m->ndf); memcpy(m->acc, a, sizeof(double) * m->nn * m->ndf);
}
void orc_get_vel_accel(void* h, double* v, double* a) {
  OrcModel* m = (OrcModel*)h;
  memcpy(v, m->vel, sizeof(double) * m->nn * m->ndf); memcpy(a, m->acc, sizeof(double) * m->nn * m->ndf);
}

/* uniaxialMaterial Steel02 / Concrete02 ; section Fiber (fibers in the order given) */
int orc_add_uniaxial(void* h, int tag, int kind, const double* p) {
  OrcModel* m = (OrcModel*)h;
  m->uni_tag = (int*)realloc(m->uni_tag, sizeof(int) * (m->nuni + 1));
  m->uni_kind = (int*)realloc(m->uni_kind, sizeof(int) * (m->nuni + 1));
  m->uni_par = (double*)realloc(m->uni_par, sizeof(double) * 12 * (m->nuni + 1));
  m->uni_tag[m->nuni] = tag; m->uni_kind[m->nuni] = kind;
  memset(m->uni_par + 12 * m->nuni, 0, 12 * sizeof(double));
  memcpy(m->uni_par + 12 * m->nuni, p, sizeof(double) * (kind == ORC_UNI_STEEL02 ? 11 : (kind == ORC_UNI_ELASTIC ? 3 : ((kind == ORC_UNI_CONCRETE01 || kind == ORC_UNI_ELASTICPP) ? 4 : 7))));
  m->nuni++; return 0;
}
int orc_add_fiber_section(void* h, int tag, int nf, const double* y, const double* A, const int* matTags) {
  OrcModel* m = (OrcModel*)h;
  m->sec = (OrcSecDef*)realloc(m->sec, sizeof(OrcSecDef) * (m->nsec + 1));
  OrcSecDef* d = &m->sec[m->nsec];
  d->tag = tag; d->nf = nf; d->z = NULL; d->GJ = 0.0; d->agg = 0;
  d->y = (double*)malloc(sizeof(double) * nf); d->A = (double*)malloc(sizeof(double) * nf); d->mat = (int*)malloc(sizeof(int) * nf);
  memcpy(d->y, y, sizeof(double) * nf); memcpy(d->A, A, sizeof(double) * nf);
  for (int i = 0; i < nf; i++) {
    d->mat[i] = -1;
    for (int j = 0; j < m->nuni; j++) if (m->uni_tag[j] == matTags[i]) d->mat[i] = j;
    if (d->mat[i] < 0) return -1;
  }
  m->nsec++; return 0;
}
/* section Aggregator tag mat1 P mat2 Mz (runtime/commands/modeling/section.cpp -> SectionAggregator.cpp:119): two
 * uniaxial materials, the first on the axial strain, the second on the curvature -- the order ForceBeamColumn2d's
 * sections carry (codes: 2 = P, 1 = Mz, SectionForceDeformation.h) */
int orc_add_section_aggregator(void* h, int tag, int n, const int* matTags, const int* codes) {
  if (n != 2 || codes[0] != 2 || codes[1] != 1) return -2;
  const double zero[2] = {0.0, 0.0}, one[2] = {1.0, 1.0};
  int rc = orc_add_fiber_section(h, tag, 2, zero, one, matTags);
  if (rc < 0) return rc;
  OrcModel* m = (OrcModel*)h;
  m->sec[m->nsec - 1].agg = 1;
  return 0;
}
/* section Fiber tag -GJ gj { fiber y z A mat ... } in a 3D model: FiberSection3d */
int orc_add_fiber_section3d(void* h, int tag, int nf, const double* y, const double* z, const double* A, const int* matTags, double GJ) {
  int rc = orc_add_fiber_section(h, tag, nf, y, A, matTags);
  if (rc < 0) return rc;
  OrcModel* m = (OrcModel*)h;
  OrcSecDef* d = &m->sec[m->nsec - 1];
  d->z = (double*)malloc(sizeof(double) * nf); memcpy(d->z, z, sizeof(double) * nf);
  d->GJ = GJ;
  return 0;
}
static int beam_update(OrcBeam* b, const double* ug, const double* dug);
static int find_mat(const OrcModel* m, int tag) { for (int i = 0; i < m->nmat; i++) if (m->mat_tag[i] == tag) return i; return -1; }

/* a fresh OrcBeam3 for element e: sections with their fibres in the initial state, transformation, zero element state */
static OrcBeam3* beam3_build(OrcModel* m, OrcEle* e, int sd, const double* par) {
  OrcBeam3* b = (OrcBeam3*)calloc(1, sizeof(OrcBeam3));
  b->nip = (int)par[0]; b->maxIters = (int)par[1]; b->tol = par[2];
  if (b->nip < 2 || b->nip > ORC_MAXSEC) return NULL;
  const OrcSecDef* d = &m->sec[sd];
  for (int i = 0; i < b->nip; i++) {
    OrcSec3* S = &b->sec[i];
    S->nf = d->nf; S->y = d->y; S->z = d->z; S->A = d->A; S->GJ = d->GJ;
    S->mat = (OrcUni*)malloc(sizeof(OrcUni) * d->nf);
    double ABar = 0.0, QzBar = 0.0, QyBar = 0.0;
    for (int f = 0; f < d->nf; f++) {
      uni_init(&S->mat[f], m->uni_kind[d->mat[f]], m->uni_par + 12 * d->mat[f]);
      S->mat[f].in_fibre = 1;
      ABar += d->A[f]; QzBar += d->y[f] * d->A[f]; QyBar += d->z[f] * d->A[f];   /* FiberSection3d::addFiber */
      S->yBar = QzBar / ABar; S->zBar = QyBar / ABar;
    }
  }
  /* par[8..13]: -jntOffset dXi dYi dZi dXj dYj dZj: the element runs between the offset ends (computeElemtLengthAndOrient) */
  double xi3[3], xj3[3];
  for (int q = 0; q < 6; q++) { b->off[q] = par[8 + q]; if (b->off[q] != 0.0) b->has_off = 1; }
  for (int q = 0; q < 3; q++) { xi3[q] = m->crd[e->node[0] * 3 + q] + b->off[q]; xj3[q] = m->crd[e->node[1] * 3 + q] + b->off[3 + q]; }
  if (crd3d_init(b, xi3, xj3, par + 3) < 0) return NULL;
  b->pdelta = (int)par[6];     /* par[6]: 0 geomTransf Linear, 1 PDelta */
  return b;
}
/* a fresh OrcBeam for element e: sections with their fibres in the initial state, transformation, zero element state */
static OrcBeam* beam2_build(OrcModel* m, OrcEle* e, int sd, const double* par) {
  OrcBeam* b = (OrcBeam*)calloc(1, sizeof(OrcBeam));
  b->nip = (int)par[0]; b->maxIters = (int)par[1]; b->tol = par[2];
  if (b->nip < 2 || b->nip > ORC_MAXSEC) return NULL;
  const OrcSecDef* d = &m->sec[sd];
  for (int i = 0; i < b->nip; i++) {
    OrcSec* S = &b->sec[i];
    S->nf = d->nf; S->y = d->y; S->A = d->A; S->agg = d->agg;
    S->mat = (OrcUni*)malloc(sizeof(OrcUni) * d->nf);
    double ABar = 0.0, QzBar = 0.0;
    for (int f = 0; f < d->nf; f++) {
      uni_init(&S->mat[f], m->uni_kind[d->mat[f]], m->uni_par + 12 * d->mat[f]);
      S->mat[f].in_fibre = d->agg ? 0 : 1;
      ABar += d->A[f]; QzBar += d->y[f] * d->A[f]; S->yBar = QzBar / ABar;   /* FiberSection2d::addFiber */
    }
    if (S->agg) { S->yBar = 0.0; S->k[0] = S->mat[0].e; S->k[3] = S->mat[1].e; }
  }
  /* LinearCrdTransf2d::computeElemtLengthAndOrient */
  double dx0 = m->crd[e->node[1] * 2] - m->crd[e->node[0] * 2], dx1 = m->crd[e->node[1] * 2 + 1] - m->crd[e->node[0] * 2 + 1];
  /* par[5..8]: -jntOffset dXi dYi dXj dYj (computeElemtLengthAndOrient adds nodeJOffset, subtracts nodeIOffset) */
  for (int q = 0; q < 4; q++) { b->off[q] = par[5 + q]; if (b->off[q] != 0.0) b->has_off = 1; }
  if (b->has_off) { dx0 += b->off[2]; dx1 += b->off[3]; dx0 -= b->off[0]; dx1 -= b->off[1]; }
  b->L = sqrt(dx0 * dx0 + dx1 * dx1);
  b->cosTheta = dx0 / b->L; b->sinTheta = dx1 / b->L;
  b->pdelta = (int)par[3] == 1; b->corot = (int)par[3] == 2; b->utrial = m->trial; b->n0 = e->node[0]; b->n1 = e->node[1];   /* par[3]: 0 geomTransf Linear, 1 PDelta, 2 Corotational */
  if (b->corot && b->has_off) return NULL;
  return b;
}
static int quad_update(OrcModel* m, OrcEle* el);
int orc_add_element(void* h, int kind, int tag, const int* nodeTags, int matTag, const double* par) {
  OrcModel* m = (OrcModel*)h;
  if (m->ne == m->ecap) { m->ecap = m->ecap ? 2 * m->ecap : 64; m->ele = (OrcEle*)realloc(m->ele, sizeof(OrcEle) * m->ecap); }
  OrcEle* e = &m->ele[m->ne];
  memset(e, 0, sizeof *e);
  e->kind = kind; e->tag = tag;
  e->nen = (kind == ORC_ELE_BRICK) ? 8 : (kind == ORC_ELE_QUAD ? 4 : 2);
  e->ndf_e = (kind == ORC_ELE_BRICK) ? 3 : (kind == ORC_ELE_QUAD ? 2 : (kind == ORC_ELE_FBC3D ? 6 : 3));
  e->nip = (kind == ORC_ELE_BRICK) ? 8 : 4;
  for (int i = 0; i < e->nen; i++) { e->node[i] = find_node(m, nodeTags[i]); if (e->node[i] < 0) return -1; }
  if (kind == ORC_ELE_FBC3D) {
    /* forceBeamColumn in a 3D model (ForceBeamColumn3d): matTag names the FiberSection3d;
     * par = nIP, maxIters, tol, vecxz[3] of `geomTransf Linear`; Lobatto integration */
    int sd = -1;
    for (int i = 0; i < m->nsec; i++) if (m->sec[i].tag == matTag) sd = i;
    if (sd < 0 || m->sec[sd].z == NULL) return -2;
    OrcBeam3* b = beam3_build(m, e, sd, par);
    if (!b) return -3;
    e->beam3 = b; e->nip = b->nip; e->mat = sd;
    memcpy(e->par, par, 16 * sizeof(double));
    double ug[12], dug[12];   /* Domain::addElement calls element->update() (Domain.cpp:391) */
    for (int a = 0; a < 2; a++) for (int j = 0; j < 6; j++) { ug[a * 6 + j] = m->trial[e->node[a] * 6 + j]; dug[a * 6 + j] = m->incr[e->node[a] * 6 + j]; }
    if (beam3_update(b, ug, dug) < 0) return -4;
    m->ne++; return 0;
  }
  if (kind == ORC_ELE_FBC2D) {
    /* forceBeamColumn: matTag names the fibre section; par = nIP, maxIters, tol.
     * Lobatto integration, Linear transformation (no offsets). */
    int sd = -1;
    for (int i = 0; i < m->nsec; i++) if (m->sec[i].tag == matTag) sd = i;
    if (sd < 0) return -2;
    OrcBeam* b = beam2_build(m, e, sd, par);
    if (!b) return -3;
    e->beam = b; e->nip = b->nip; e->mat = sd;
    memcpy(e->par, par, 16 * sizeof(double));
    /* Domain::addElement calls element->update() (Domain.cpp:391) */
    double ug[6], dug[6];
    for (int a = 0; a < 2; a++) for (int j = 0; j < 3; j++) { ug[a * 3 + j] = m->trial[e->node[a] * 3 + j]; dug[a * 3 + j] = m->incr[e->node[a] * 3 + j]; }
    if (beam_update(b, ug, dug) < 0) return -4;
    m->ne++; return 0;
  }
  e->mat = find_mat(m, matTag); if (e->mat < 0) return -2;
  memset(e->par, 0, sizeof e->par); memcpy(e->par, par, 8 * sizeof(double));
  int type = (kind == ORC_ELE_BRICK) ? ORC_ND_3D : ((int)par[1] == 1 ? ORC_ND_PLANE_STRESS : ORC_ND_PLANE_STRAIN);
  for (int i = 0; i < e->nip; i++) gp_init(&e->gp[i], m->mat_kind[e->mat], type, m->mat_par + 8 * e->mat);
  /* Domain::addElement calls element->update() (Domain.cpp:391): J2PlaneStress condenses its tangent there */
  if (type == ORC_ND_PLANE_STRESS && m->mat_kind[e->mat] == ORC_MAT_J2) quad_update(m, e);
  m->ne++; return 0;
}
int orc_add_load(void* h, int nodeTag, const double* v) {
  OrcModel* m = (OrcModel*)h; int n = find_node(m, nodeTag); if (n < 0) return -1;
  for (int i = 0; i < m->ndf; i++) m->load[n * m->ndf + i] += v[i];
  return 0;
}

/* loadConst (Domain::setLoadConstant): the loads applied so far keep their current factor; the next pattern starts empty.
 * orc_add_load then fills that pattern. */
int orc_load_const(void* h) {
  OrcModel* m = (OrcModel*)h;
  const size_t n = (size_t)m->nn * m->ndf;
  if (!m->cload) m->cload = (double*)calloc(n ? n : 1, sizeof(double));
  for (size_t i = 0; i < n; i++) { m->cload[i] = m->cload[i] + m->load[i] * m->lambda; m->load[i] = 0.0; }
  if (!m->ele_loads_const) { m->ele_loads_const = 1; m->ele_lambda = m->lambda; }
  return 0;
}

static int cmp_ele(const void* a, const void* b) { return ((const OrcEle*)a)->tag - ((const OrcEle*)b)->tag; }

/* sorted-unique insert, restating ID::insert (SRC/matrix/ID.cpp:512-575) */
typedef struct { int* d; int n, cap; } IntSet;
static int iset_insert(IntSet* s, int x) {
  int left = 0, right = s->n - 1, middle;
  while (left <= right) { middle = (left + right) / 2; if (x == s->d[middle]) return 1; if (x > s->d[middle]) left = middle + 1; else right = middle - 1; }
  if (s->n == s->cap) { s->cap = (s->n + 1) * 2; s->d = (int*)realloc(s->d, sizeof(int) * s->cap); }
  for (int i = s->n; i > left; i--) s->d[i] = s->d[i - 1];
  s->d[left] = x; s->n++; return 0;
}

/* PlainHandler::handle (PlainHandler.cpp:60-250) -> initial ids (-2 free, -1 fixed);
 * numberer 0: PlainNumberer::numberDOF (PlainNumberer.cpp:73-157)
 * numberer 1: DOF_Numberer::numberDOF (DOF_Numberer.cpp:92-205) over
 *             AnalysisModel::getDOFGroupGraph (AnalysisModel.cpp:355-400) with
 *             RCM::number, GPS off (RCM.cpp:66-276; numberer.cpp:46 builds RCM(false))
 * then AnalysisModel::getDOFGraph (AnalysisModel.cpp:286-351) and
 * SparseGenColLinSOE::setSize (SparseGenColLinSOE.cpp:161-261) /
 * SparseGenRowLinSOE::setSize (SparseGenRowLinSOE.cpp:127-220). */
int orc_setup(void* h, int numberer, int soe_kind) {
  OrcModel* m = (OrcModel*)h;
  int nn = m->nn, ndf = m->ndf;
  qsort(m->ele, m->ne, sizeof(OrcEle), cmp_ele); /* Domain element map iterates by tag */
  m->id = (int*)malloc(sizeof(int) * nn * ndf);
  for (int i = 0; i < nn * ndf; i++) m->id[i] = (m->fixed[i] || i % ndf >= m->node_ndf[i / ndf]) ? -1 : -2;
  /* FourNodeQuad.cpp:133-139 (and Brick / ForceBeamColumn alike): the element's nodes must carry its dofs per node */
  for (int e = 0; e < m->ne; e++)
    for (int i = 0; i < m->ele[e].nen; i++) if (m->node_ndf[m->ele[e].node[i]] != m->ele[e].ndf_e) return -7;
  for (int i = 0; i < m->nmp; i++) {       /* PlainHandler: -4 unless the dof is already constrained (then a warning) */
    int* idc = &m->id[m->mp[3 * i + 1] * ndf + m->mp[3 * i + 2]];
    if (*idc == -2) *idc = -4;
  }

  int* order = (int*)malloc(sizeof(int) * nn);
  if (numberer == 0) {
    for (int i = 0; i < nn; i++) order[i] = i;
  } else {
    /* DOF_Group graph: vertex per node (DOF_Group tag = position in tag order) */
    IntSet* adj = (IntSet*)calloc(nn, sizeof(IntSet));
    for (int e = 0; e < m->ne; e++) {
      OrcEle* el = &m->ele[e];
      for (int i = 0; i < el->nen; i++)
        for (int j = 0; j < el->nen; j++)
          if (i != j && el->node[i] != el->node[j]) {
            /* Graph::addEdge (Graph.cpp:178-213): insert in both, skip when present */
            if (iset_insert(&adj[el->node[i]], el->node[j]) == 0) iset_insert(&adj[el->node[j]], el->node[i]);
          }
    }
    int* tmp = (int*)malloc(sizeof(int) * nn);
    for (int i = 0; i < nn; i++) tmp[i] = -1;
    int iter = 0;                       /* vertexIter4 */
    int currentMark = nn - 1, nextMark = currentMark - 1;
    order[currentMark] = 0; tmp[0] = currentMark;   /* first vertex from the iterator */
    iter = 0;
    while (nextMark >= 0) {
      int v = order[currentMark];
      for (int a = 0; a < adj[v].n; a++) {
        int w = adj[v].d[a];
        if (tmp[w] == -1) { tmp[w] = nextMark; order[nextMark--] = w; }
      }
      currentMark--;
      if (currentMark == nextMark && currentMark >= 0) {
        while (iter < nn && tmp[iter] != -1) iter++;
        nextMark--;
        tmp[iter] = currentMark; order[currentMark] = iter; iter++;
      }
    }
    for (int i = 0; i < nn; i++) free(adj[i].d);
    free(adj); free(tmp);
  }
  int eqn = 0;
  for (int i = 0; i < nn; i++) {
    int n = order[i];
    for (int j = 0; j < ndf; j++) if (m->id[n * ndf + j] == -2) m->id[n * ndf + j] = eqn++;
  }
  free(order);
  m->neq = eqn;
  for (int i = 0; i < m->nmp; i++) {       /* the numberer's last pass: -4 -> the retained dof's id */
    int* idc = &m->id[m->mp[3 * i + 1] * ndf + m->mp[3 * i + 2]];
    if (*idc == -4) *idc = m->id[m->mp[3 * i] * ndf + m->mp[3 * i + 2]];
  }

  /* DOF graph: per equation a sorted adjacency (Vertex::addEdge skips self) */
  IntSet* adj = (IntSet*)calloc(eqn > 0 ? eqn : 1, sizeof(IntSet));
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    int ids[24], n = 0;
    for (int i = 0; i < el->nen; i++) for (int j = 0; j < el->ndf_e; j++) ids[n++] = m->id[el->node[i] * ndf + j];
    for (int i = 0; i < n; i++) if (ids[i] >= 0)
      for (int j = i + 1; j < n; j++) if (ids[j] >= 0 && ids[j] != ids[i])
        if (iset_insert(&adj[ids[i]], ids[j]) == 0) iset_insert(&adj[ids[j]], ids[i]);
  }
  m->soe_kind = soe_kind;
  m->ptr = (int*)malloc(sizeof(int) * (eqn + 1));
  int nnz = 0; for (int a = 0; a < eqn; a++) nnz += adj[a].n + 1;
  m->nnz = nnz;
  m->idx = (int*)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
  m->ptr[0] = 0;
  int startLoc = 0, lastLoc = 0;
  for (int a = 0; a < eqn; a++) {
    m->idx[lastLoc++] = a;                       /* "place diag in first" */
    for (int i = 0; i < adj[a].n; i++) {
      int row = adj[a].d[i], found = 0;
      for (int j = startLoc; j < lastLoc; j++)
        if (m->idx[j] > row) {
          for (int k = lastLoc; k > j; k--) m->idx[k] = m->idx[k - 1];
          m->idx[j] = row; found = 1; j = lastLoc;
        }
      if (!found) m->idx[lastLoc] = row;
      lastLoc++;
    }
    m->ptr[a + 1] = lastLoc; startLoc = lastLoc;
  }
  for (int a = 0; a < eqn; a++) free(adj[a].d);
  free(adj);
  m->A = (double*)calloc(nnz > 0 ? nnz : 1, sizeof(double));
  m->B = (double*)calloc(eqn > 0 ? eqn : 1, sizeof(double));
  return eqn;
}

int orc_num_eqn(void* h) { return ((OrcModel*)h)->neq; }
int orc_nnz(void* h) { return ((OrcModel*)h)->nnz; }
void orc_get_ids(void* h, int* ids) { OrcModel* m = (OrcModel*)h; memcpy(ids, m->id, sizeof(int) * m->nn * m->ndf); }
void orc_get_csr(void* h, int* ptr, int* idx) { OrcModel* m = (OrcModel*)h; memcpy(ptr, m->ptr, sizeof(int) * (m->neq + 1)); memcpy(idx, m->idx, sizeof(int) * m->nnz); }
int orc_num_ele(void* h) { return ((OrcModel*)h)->ne; }

/* FE_Element::setID (FE_Element.cpp:209-231): element dof -> equation */
static int ele_ids(const OrcModel* m, const OrcEle* el, int* ids) {
  int n = 0;
  for (int i = 0; i < el->nen; i++) for (int j = 0; j < el->ndf_e; j++) ids[n++] = m->id[el->node[i] * m->ndf + j];
  return n;
}
/* FE ids in FE_Element order, [ne][stride] */
void orc_fe_ids(void* h, int* tags, int* ids, int stride) {
  OrcModel* m = (OrcModel*)h;
  for (int e = 0; e < m->ne; e++) { int t[32]; int n = ele_ids(m, &m->ele[e], t); tags[e] = m->ele[e].tag; for (int i = 0; i < n && i < stride; i++) ids[(size_t)e * stride + i] = t[i]; }
}

/* scatter map: position in A where addA puts m(i,j); -1 if dropped.
 * SparseGenColLinSOE::addA (SparseGenColLinSOE.cpp:264-330): column id(i), row id(j) gets m(j,i)
 * SparseGenRowLinSOE::addA (SparseGenRowLinSOE.cpp:224-282): row id(i), column id(j) gets m(i,j)
 * map[e][i*nd+j] is for element-matrix entry (i,j) in both cases. */
static int soe_find(const OrcModel* m, int major, int minor) {
  for (int k = m->ptr[major]; k < m->ptr[major + 1]; k++) if (m->idx[k] == minor) return k;
  return -1;
}
/* BandGenLinSOE::setSize (bandGEN/BandGenLinSOE.cpp:116-147) / ProfileSPDLinSOE::setSize (profileSPD/ProfileSPDLinSOE.cpp:
 * 115-168) over the DOF graph (vertex = equation, adjacency = the other equations of its column).  The model must have
 * been set up with soe_kind 0; returns the length of A.  kind 0 goes back to the compressed columns. */
long long orc_set_store(void* h, int kind) {
  OrcModel* m = (OrcModel*)h;
  free(m->iDiagLoc); free(m->Astore); m->iDiagLoc = NULL; m->Astore = NULL; m->store_kind = 0; m->astore_size = m->nnz;
  if (kind != 2 && kind != 3) return m->nnz;
  if (m->soe_kind != 0) return -1;
  const int size = m->neq;
  if (kind == 2) {
    int numSubD = 0, numSuperD = 0;
    for (int v = 0; v < size; v++)
      for (int k = m->ptr[v]; k < m->ptr[v + 1]; k++) {
        int otherNum = m->idx[k];
        if (otherNum == v) continue;            /* the adjacency list does not hold the vertex itself */
        int diff = v - otherNum;
        if (diff > 0) { if (diff > numSuperD) numSuperD = diff; }
        else if (diff < numSubD) numSubD = diff;
      }
    numSubD *= -1;
    m->numSubD = numSubD; m->numSuperD = numSuperD;
    m->astore_size = (long long)size * (2 * numSubD + numSuperD + 1);
  } else {
    m->iDiagLoc = (int*)calloc(size > 0 ? size : 1, sizeof(int));
    for (int v = 0; v < size; v++)
      for (int k = m->ptr[v]; k < m->ptr[v + 1]; k++) {
        int diff = v - m->idx[k];
        if (diff > 0 && m->iDiagLoc[v] < diff) m->iDiagLoc[v] = diff;
      }
    if (size > 0) m->iDiagLoc[0] = 1;          /* NOTE FORTRAN ARRAY LOCATION */
    for (int j = 1; j < size; j++) m->iDiagLoc[j] = m->iDiagLoc[j] + 1 + m->iDiagLoc[j - 1];
    m->profileSize = size > 0 ? m->iDiagLoc[size - 1] : 0;
    m->astore_size = m->profileSize;
  }
  m->store_kind = kind;
  m->Astore = (double*)calloc(m->astore_size > 0 ? m->astore_size : 1, sizeof(double));
  return m->astore_size;
}
void orc_get_band(void* h, int* out) { OrcModel* m = (OrcModel*)h; out[0] = m->numSubD; out[1] = m->numSuperD; }
void orc_get_profile(void* h, int* iDiagLoc) { OrcModel* m = (OrcModel*)h; memcpy(iDiagLoc, m->iDiagLoc, sizeof(int) * m->neq); }
/* where BandGenLinSOE::addA (BandGenLinSOE.cpp:208-249) / ProfileSPDLinSOE::addA (ProfileSPDLinSOE.cpp:214-243) put the
 * entry of column col, row row; -1 when it is dropped */
static long long store_loc(const OrcModel* m, int row, int col) {
  if (m->store_kind == 2) {
    const int ldA = 2 * m->numSubD + m->numSuperD + 1;
    long long colii = (long long)col * ldA + m->numSubD + m->numSuperD;
    int diff = col - row;
    if (diff > 0) return diff <= m->numSuperD ? colii - diff : -1;
    diff *= -1;
    return diff <= m->numSubD ? colii + diff : -1;
  }
  int minColRow = col == 0 ? 0 : col - (m->iDiagLoc[col] - m->iDiagLoc[col - 1]) + 1;
  if (row <= col && row >= minColRow) return (long long)m->iDiagLoc[col] - 1 + (row - col);
  return -1;
}
void orc_scatter_map(void* h, int e, int* map) {
  OrcModel* m = (OrcModel*)h; int ids[32]; int n = ele_ids(m, &m->ele[e], ids);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
    int r = ids[i], c = ids[j], k = -1;
    /* column-oriented addA: column id(j'), row id(i') receives m(i', j') -- entry (i, j) goes to column c = id(j), row r = id(i) */
    if (r >= 0 && c >= 0) k = m->store_kind ? (int)store_loc(m, r, c) : ((m->soe_kind == 1) ? soe_find(m, r, c) : soe_find(m, c, r));
    map[i * n + j] = k;
  }
}

/* ---- element state determination ---------------------------------------- */


/* Brick::update, Brick.cpp:718-840 */
static int brick_update(OrcModel* m, OrcEle* el) {
  double xl[3][8];
  for (int i = 0; i < 8; i++) for (int d = 0; d < 3; d++) xl[d][i] = m->crd[el->node[i] * 3 + d];
  const double one_over_root3 = 1.0 / sqrt(3.0);
  const double sg[2] = { -one_over_root3, one_over_root3 };
  int count = 0, rc = 0;

  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
    double gp[3] = { sg[i], sg[j], sg[k] }, xsj, shp[4][8], strain[6] = {0, 0, 0, 0, 0, 0};
    shp3d(gp, &xsj, shp, xl);
    for (int n = 0; n < 8; n++) {
      const double* ul = m->trial + (size_t)el->node[n] * m->ndf;
      double b00 = shp[0][n], b11 = shp[1][n], b22 = shp[2][n], b30 = shp[1][n], b31 = shp[0][n],
             b41 = shp[2][n], b42 = shp[1][n], b50 = shp[2][n], b52 = shp[0][n];
      strain[0] += b00 * ul[0];
      strain[1] += b11 * ul[1];
      strain[2] += b22 * ul[2];
      strain[3] += b30 * ul[0] + b31 * ul[1];
      strain[4] += b41 * ul[1] + b42 * ul[2];
      strain[5] += b50 * ul[0] + b52 * ul[2];
    }
    if (gp_set_trial_strain(&el->gp[count], strain) < 0) rc = -1;
    count++;
  }
  return rc;
}

/* Brick::formResidAndTangent, Brick.cpp:843-1022; K row-major [24][24], R[24].
 * The tangent products keep Matrix::addMatrixProduct's j,k,i loop order
 * (Matrix.cpp:693-720) so the floating point sums match the reference. */
static void brick_form(OrcModel* m, OrcEle* el, int tang_flag, double* K, double* R) {
  double xl[3][8];
  for (int i = 0; i < 8; i++) for (int d = 0; d < 3; d++) xl[d][i] = m->crd[el->node[i] * 3 + d];
  const double one_over_root3 = 1.0 / sqrt(3.0);
  const double sg[2] = { -one_over_root3, one_over_root3 };
  double Shape[8][4][8], dvol[8];
  int count = 0;
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
    double gp[3] = { sg[i], sg[j], sg[k] }, xsj;
    shp3d(gp, &xsj, Shape[count], xl);
    dvol[count] = 1.0 * xsj;
    count++;
  }
  if (K) memset(K, 0, 576 * sizeof(double));
  memset(R, 0, 24 * sizeof(double));
  for (int g = 0; g < 8; g++) {
    double (*shp)[8] = Shape[g];
    double stress[6], dd[36];
    gp_get_stress(&el->gp[g], stress);
    for (int i = 0; i < 6; i++) stress[i] *= dvol[g];
    if (tang_flag) {   /* 2: Brick::getInitialStiff (Brick.cpp:324), the same loops on getInitialTangent() */
      if (tang_flag == 2) gp_get_initial_tangent(&el->gp[g], dd); else gp_get_tangent(&el->gp[g], dd);
      for (int i = 0; i < 36; i++) dd[i] *= dvol[g];
    }
    int jj = 0;
    for (int j = 0; j < 8; j++) {
      double b00 = shp[0][j], b11 = shp[1][j], b22 = shp[2][j], b30 = shp[1][j], b31 = shp[0][j],
             b41 = shp[2][j], b42 = shp[1][j], b50 = shp[2][j], b52 = shp[0][j];
      double residJ[3];
      residJ[0] = b00 * stress[0] + b30 * stress[3] + b50 * stress[5];
      residJ[1] = b11 * stress[1] + b31 * stress[3] + b41 * stress[4];
      residJ[2] = b22 * stress[2] + b42 * stress[4] + b52 * stress[5];
      for (int p = 0; p < 3; p++) {
        R[jj + p] += residJ[p];
        R[jj + p] -= dvol[g] * el->par[p] * shp[3][j];
      }
      if (tang_flag) {
        /* BJ (6x3), BJtran (3x6) */
        double BJ[6][3] = {{shp[0][j], 0, 0}, {0, shp[1][j], 0}, {0, 0, shp[2][j]},
                           {shp[1][j], shp[0][j], 0}, {0, shp[2][j], shp[1][j]}, {shp[2][j], 0, shp[0][j]}};
        double BJtranD[3][6];
        /* BJtranD = BJtran * dd : for col j2, for k2, tmp = dd(k2,j2); for i2: out(i2,j2) += BJtran(i2,k2)*tmp */
        for (int j2 = 0; j2 < 6; j2++) {
          for (int i2 = 0; i2 < 3; i2++) BJtranD[i2][j2] = 0.0;
          for (int k2 = 0; k2 < 6; k2++) {
            double tmp = dd[k2 * 6 + j2] * 1.0;
            for (int i2 = 0; i2 < 3; i2++) BJtranD[i2][j2] += BJ[k2][i2] * tmp;
          }
        }
        int kk = 0;
        for (int k = 0; k < 8; k++) {
          double BK[6][3] = {{shp[0][k], 0, 0}, {0, shp[1][k], 0}, {0, 0, shp[2][k]},
                             {shp[1][k], shp[0][k], 0}, {0, shp[2][k], shp[1][k]}, {shp[2][k], 0, shp[0][k]}};
          double stiffJK[3][3];
          for (int j2 = 0; j2 < 3; j2++) {
            for (int i2 = 0; i2 < 3; i2++) stiffJK[i2][j2] = 0.0;
            for (int k2 = 0; k2 < 6; k2++) {
              double tmp = BK[k2][j2] * 1.0;
              for (int i2 = 0; i2 < 3; i2++) stiffJK[i2][j2] += BJtranD[i2][k2] * tmp;
            }
          }
          for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++) K[(jj + p) * 24 + kk + q] += stiffJK[p][q];
          kk += 3;
        }
      }
      jj += 3;
    }
  }
}

/* FourNodeQuad::shapeFunction, FourNodeQuad.cpp:1128-1196 */
static double quad_shape(const OrcModel* m, const OrcEle* el, double xi, double eta, double shp[3][4]) {
  const double* c1 = m->crd + el->node[0] * 2; const double* c2 = m->crd + el->node[1] * 2;
  const double* c3 = m->crd + el->node[2] * 2; const double* c4 = m->crd + el->node[3] * 2;
  double oneMinuseta = 1.0 - eta, onePluseta = 1.0 + eta, oneMinusxi = 1.0 - xi, onePlusxi = 1.0 + xi;
  shp[2][0] = 0.25 * oneMinusxi * oneMinuseta;
  shp[2][1] = 0.25 * onePlusxi * oneMinuseta;
  shp[2][2] = 0.25 * onePlusxi * onePluseta;
  shp[2][3] = 0.25 * oneMinusxi * onePluseta;
  double J[2][2], L[2][2];
  J[0][0] = 0.25 * (-c1[0] * oneMinuseta + c2[0] * oneMinuseta + c3[0] * (onePluseta) - c4[0] * (onePluseta));
  J[0][1] = 0.25 * (-c1[0] * oneMinusxi - c2[0] * onePlusxi + c3[0] * onePlusxi + c4[0] * oneMinusxi);
  J[1][0] = 0.25 * (-c1[1] * oneMinuseta + c2[1] * oneMinuseta + c3[1] * onePluseta - c4[1] * onePluseta);
  J[1][1] = 0.25 * (-c1[1] * oneMinusxi - c2[1] * onePlusxi + c3[1] * onePlusxi + c4[1] * oneMinusxi);
  double detJ = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  double oneOverdetJ = 1.0 / detJ;
  L[0][0] = J[1][1] * oneOverdetJ; L[1][0] = -J[0][1] * oneOverdetJ;
  L[0][1] = -J[1][0] * oneOverdetJ; L[1][1] = J[0][0] * oneOverdetJ;
  double L00 = 0.25 * L[0][0], L10 = 0.25 * L[1][0], L01 = 0.25 * L[0][1], L11 = 0.25 * L[1][1];
  double L00oneMinuseta = L00 * oneMinuseta, L00onePluseta = L00 * onePluseta;
  double L01oneMinusxi = L01 * oneMinusxi, L01onePlusxi = L01 * onePlusxi;
  double L10oneMinuseta = L10 * oneMinuseta, L10onePluseta = L10 * onePluseta;
  double L11oneMinusxi = L11 * oneMinusxi, L11onePlusxi = L11 * onePlusxi;
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
/* integration points: FourNodeQuad.h pts/wts (Gauss 2x2, counter-clockwise) */
static const double quad_pts[4][2] = {
  {-0.577350269189626, -0.577350269189626}, { 0.577350269189626, -0.577350269189626},
  { 0.577350269189626,  0.577350269189626}, {-0.577350269189626,  0.577350269189626}};
static const double quad_wts[4] = {1.0, 1.0, 1.0, 1.0};

/* FourNodeQuad::update, FourNodeQuad.cpp:190-222 */
static int quad_update(OrcModel* m, OrcEle* el) {
  double u[2][4]; int ret = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 2; j++) u[j][i] = m->trial[(size_t)el->node[i] * m->ndf + j];
  for (int i = 0; i < 4; i++) {
    double shp[3][4]; quad_shape(m, el, quad_pts[i][0], quad_pts[i][1], shp);
    double eps[3] = {0, 0, 0};
    for (int beta = 0; beta < 4; beta++) {
      eps[0] += shp[0][beta] * u[0][beta];
      eps[1] += shp[1][beta] * u[1][beta];
      eps[2] += shp[0][beta] * u[1][beta] + shp[1][beta] * u[0][beta];
    }
    ret += gp_set_trial_strain(&el->gp[i], eps);
  }
  return ret;
}
/* FourNodeQuad::getTangentStiff (FourNodeQuad.cpp:226-281), getResistingForce (507-553);
 * with the surface pressure load; Q (element loads) is zero in the hot path models. */
static void quad_form(OrcModel* m, OrcEle* el, int tang_flag, double* K, double* P) {
  double thickness = el->par[0];
  if (tang_flag) {
    memset(K, 0, 64 * sizeof(double));
    for (int i = 0; i < 4; i++) {
      double shp[3][4]; double dvol = quad_shape(m, el, quad_pts[i][0], quad_pts[i][1], shp);
      dvol *= (thickness * quad_wts[i]);
      double D[9];
      if (tang_flag == 2) gp_get_initial_tangent(&el->gp[i], D); else gp_get_tangent(&el->gp[i], D);   /* 2: getInitialStiff */
      const double D00 = D[0], D01 = D[1], D02 = D[2], D10 = D[3], D11 = D[4], D12 = D[5], D20 = D[6], D21 = D[7], D22 = D[8];
      double DB[3][2];
      for (int alpha = 0, ia = 0; alpha < 4; alpha++, ia += 2)
        for (int beta = 0, ib = 0; beta < 4; beta++, ib += 2) {
          DB[0][0] = dvol * (D00 * shp[0][beta] + D02 * shp[1][beta]);
          DB[1][0] = dvol * (D10 * shp[0][beta] + D12 * shp[1][beta]);
          DB[2][0] = dvol * (D20 * shp[0][beta] + D22 * shp[1][beta]);
          DB[0][1] = dvol * (D01 * shp[1][beta] + D02 * shp[0][beta]);
          DB[1][1] = dvol * (D11 * shp[1][beta] + D12 * shp[0][beta]);
          DB[2][1] = dvol * (D21 * shp[1][beta] + D22 * shp[0][beta]);
          K[ia * 8 + ib]           += shp[0][alpha] * DB[0][0] + shp[1][alpha] * DB[2][0];
          K[ia * 8 + ib + 1]       += shp[0][alpha] * DB[0][1] + shp[1][alpha] * DB[2][1];
          K[(ia + 1) * 8 + ib]     += shp[1][alpha] * DB[1][0] + shp[0][alpha] * DB[2][0];
          K[(ia + 1) * 8 + ib + 1] += shp[1][alpha] * DB[1][1] + shp[0][alpha] * DB[2][1];
        }
    }
  }
  memset(P, 0, 8 * sizeof(double));
  for (int i = 0; i < 4; i++) {
    double shp[3][4]; double dvol = quad_shape(m, el, quad_pts[i][0], quad_pts[i][1], shp);
    dvol *= (thickness * quad_wts[i]);
    double sigma[3]; gp_get_stress(&el->gp[i], sigma);
    for (int alpha = 0, ia = 0; alpha < 4; alpha++, ia += 2) {
      P[ia]     += dvol * (shp[0][alpha] * sigma[0] + shp[1][alpha] * sigma[2]);
      P[ia + 1] += dvol * (shp[1][alpha] * sigma[1] + shp[0][alpha] * sigma[2]);
      P[ia]     -= dvol * (shp[2][alpha] * el->par[4]);
      P[ia + 1] -= dvol * (shp[2][alpha] * el->par[5]);
    }
  }
  /* surface pressure: FourNodeQuad::setPressureLoadAtNodes (FourNodeQuad.cpp:1206-1264), P = P - pressureLoad (:542-546) */
  if (el->par[2] != 0.0) {
    double pl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double po2 = el->par[2] / 2.0;
    for (int a = 0; a < 4; a++) {
      int b = (a + 1) & 3;
      const double* xa = m->crd + (size_t)el->node[a] * m->ndm; const double* xb = m->crd + (size_t)el->node[b] * m->ndm;
      double dx = xb[0] - xa[0], dy = xb[1] - xa[1];
      pl[2 * a] += po2 * dy; pl[2 * b] += po2 * dy;
      pl[2 * a + 1] += po2 * -dx; pl[2 * b + 1] += po2 * -dx;
    }
    for (int i = 0; i < 8; i++) P[i] += pl[i] * -1.0;
  }
}

/* Node::setTrialDisp for all nodes + Domain::update -> Element::update */
static int ele_update(OrcModel* m, OrcEle* el) {
  if (el->kind == ORC_ELE_BRICK) return brick_update(m, el);
  if (el->kind == ORC_ELE_QUAD) return quad_update(m, el);
  if (el->kind == ORC_ELE_FBC3D) {
    double ug[12], dug[12];
    for (int a = 0; a < 2; a++) for (int j = 0; j < 6; j++) { ug[a * 6 + j] = m->trial[el->node[a] * 6 + j]; dug[a * 6 + j] = m->incr[el->node[a] * 6 + j]; }
    return beam3_update(el->beam3, ug, dug);
  }
  double ug[6], dug[6];
  for (int a = 0; a < 2; a++) for (int j = 0; j < 3; j++) { ug[a * 3 + j] = m->trial[el->node[a] * 3 + j]; dug[a * 3 + j] = m->incr[el->node[a] * 3 + j]; }
  return beam_update(el->beam, ug, dug);
}
/* Node::setTrialDisp (Node.cpp: incrDeltaDisp = new - trial) + Domain::update */
int orc_set_trial_disp(void* h, const double* u) {
  OrcModel* m = (OrcModel*)h; int rc = 0;
  for (int i = 0; i < m->nn * m->ndf; i++) { m->incr[i] = u[i] - m->trial[i]; m->trial[i] = u[i]; }
  for (int e = 0; e < m->ne; e++) rc |= ele_update(m, &m->ele[e]);
  return rc;
}
/* Newmark::update (Newmark.cpp:411-458): U += cu*dU, Udot += cv*dU, Udotdot += ca*dU by equation,
 * AnalysisModel::setResponse -> Node::setTrialDisp/Vel/Accel, then updateDomain */
int orc_incr_response(void* h, const double* dU, double cu, double cv, double ca) {
  OrcModel* m = (OrcModel*)h; int rc = 0;
  for (int i = 0; i < m->nn * m->ndf; i++) {
    int r = m->id[i];
    if (r < 0) { m->incr[i] = 0.0; continue; }
    double d = dU[r];
    double un = (cu == 1.0) ? m->trial[i] + d : m->trial[i] + d * cu;
    m->incr[i] = un - m->trial[i]; m->trial[i] = un;
    m->vel[i] += d * cv; m->acc[i] += d * ca;
  }
  for (int e = 0; e < m->ne; e++) rc |= ele_update(m, &m->ele[e]);
  return rc;
}
/* Domain::applyLoad(pseudoTime): zeroLoad on every element, then LoadPattern::applyLoad -> NodalLoad::applyLoad and
 * ElementalLoad::applyLoad -> Element::addLoad(load, loadFactor) (domain/domain/Domain.cpp applyLoad; Linear series) */
void orc_apply_load(void* h, double lambda) {
  OrcModel* m = (OrcModel*)h;
  m->lambda = lambda;
  const double lf = m->ele_loads_const ? m->ele_lambda : lambda;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->kind == ORC_ELE_FBC2D && (el->beam->has_load || el->beam->has_point || el->beam->has_partial)) { el->beam->numEleLoads = el->beam->has_load + el->beam->has_point + el->beam->has_partial; el->beam->loadFactor = lf; }
    if (el->kind == ORC_ELE_FBC3D && (el->beam3->has_load || el->beam3->has_point || el->beam3->has_partial)) { el->beam3->numEleLoads = el->beam3->has_load + el->beam3->has_point + el->beam3->has_partial; el->beam3->loadFactor = lf; }
  }
}
/* `eleLoad -ele tag -type -beamUniform wy [wz] wa` in the model's Linear pattern (Beam2dUniformLoad / Beam3dUniformLoad) */
int orc_add_beam_uniform_load(void* h, int ele_tag, double wy, double wz, double wa) {
  OrcModel* m = (OrcModel*)h;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->tag != ele_tag) continue;
    /* (a further uniform load on the same element: the element loops over its loads and adds their terms, all at the
     *  pattern's factor -- kept here as one load with the summed intensities) */
    if (el->kind == ORC_ELE_FBC2D) { el->beam->has_load = 1; el->beam->w[0] += wy; el->beam->w[2] += wa; return 0; }
    if (el->kind == ORC_ELE_FBC3D) { el->beam3->has_load = 1; el->beam3->w[0] += wy; el->beam3->w[1] += wz; el->beam3->w[2] += wa; return 0; }
    return -3;
  }
  return -1;
}

/* `element forceBeamColumn ... -integration Legendre | Radau | NewtonCotes | Trapezoidal | ...`: the nip section locations
 * and weights (fractions of L) that the element's BeamIntegration returns, in place of the Lobatto tables */
int orc_set_beam_integration(void* h, int ele_tag, int nip, const double* xi, const double* wt) {
  OrcModel* m = (OrcModel*)h;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->tag != ele_tag) continue;
    /* the element was built (and updated once, Domain::addElement) with the Lobatto tables: build it again with its rule
     * (part of the element's definition: call this before any element load is added) */
    if (el->kind == ORC_ELE_FBC2D) {
      if (nip != el->beam->nip) return -2;
      for (int i = 0; i < el->beam->nip; i++) free(el->beam->sec[i].mat);
      free(el->beam);
      el->beam = beam2_build(m, el, el->mat, el->par);
      el->beam->user_rule = 1; memcpy(el->beam->rxi, xi, sizeof(double) * nip); memcpy(el->beam->rwt, wt, sizeof(double) * nip);
      double ug[6], dug[6];
      for (int a = 0; a < 2; a++) for (int j = 0; j < 3; j++) { ug[a * 3 + j] = m->trial[el->node[a] * 3 + j]; dug[a * 3 + j] = m->incr[el->node[a] * 3 + j]; }
      return beam_update(el->beam, ug, dug) < 0 ? -4 : 0;
    }
    if (el->kind == ORC_ELE_FBC3D) {
      if (nip != el->beam3->nip) return -2;
      for (int i = 0; i < el->beam3->nip; i++) free(el->beam3->sec[i].mat);
      free(el->beam3);
      el->beam3 = beam3_build(m, el, el->mat, el->par);
      el->beam3->user_rule = 1; memcpy(el->beam3->rxi, xi, sizeof(double) * nip); memcpy(el->beam3->rwt, wt, sizeof(double) * nip);
      double ug[12], dug[12];
      for (int a = 0; a < 2; a++) for (int j = 0; j < 6; j++) { ug[a * 6 + j] = m->trial[el->node[a] * 6 + j]; dug[a * 6 + j] = m->incr[el->node[a] * 6 + j]; }
      return beam3_update(el->beam3, ug, dug) < 0 ? -4 : 0;
    }
    return -3;
  }
  return -1;
}

/* `eleLoad -ele tag -type -beamPoint Py [Pz] xL [N]` in the model's Linear pattern (Beam2dPointLoad / Beam3dPointLoad); a
 * load with xL outside [0, 1] is ignored by the element (ForceBeamColumn2d.cpp:447) and is not kept */
int orc_add_beam_point_load(void* h, int ele_tag, double Py, double Pz, double N, double aOverL) {
  OrcModel* m = (OrcModel*)h;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->tag != ele_tag) continue;
    if (aOverL < 0.0 || aOverL > 1.0) return 0;
    if (el->kind == ORC_ELE_FBC2D) { if (el->beam->has_point) return -2; el->beam->has_point = 1; el->beam->pt[0] = Py; el->beam->pt[1] = 0.0; el->beam->pt[2] = N; el->beam->pt[3] = aOverL; return 0; }
    if (el->kind == ORC_ELE_FBC3D) { if (el->beam3->has_point) return -2; el->beam3->has_point = 1; el->beam3->pt[0] = Py; el->beam3->pt[1] = Pz; el->beam3->pt[2] = N; el->beam3->pt[3] = aOverL; return 0; }
    return -3;
  }
  return -1;
}

/* `eleLoad -beamUniform` over part of an element (Beam2d/3dPartialUniformLoad): q[8] = wy_a, wy_b, wAxial_a, wAxial_b, aOverL, bOverL, wz_a, wz_b (3D) */
int orc_add_beam_partial_load(void* h, int ele_tag, const double* q) {
  OrcModel* m = (OrcModel*)h;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->tag != ele_tag) continue;
    if (el->kind == ORC_ELE_FBC3D) {      /* q[0..7] = wy_a, wy_b, wAxial_a, wAxial_b, aOverL, bOverL, wz_a, wz_b */
      if (el->beam3->has_partial) return -2;
      el->beam3->has_partial = 1; memcpy(el->beam3->pq, q, sizeof el->beam3->pq);
      return 0;
    }
    if (el->kind != ORC_ELE_FBC2D) return -3;
    if (el->beam->has_partial) return -2;
    el->beam->has_partial = 1; memcpy(el->beam->pq, q, sizeof el->beam->pq);
    return 0;
  }
  return -1;
}

/* Element::getTangentStiff / getResistingForce of FE element e (row-major) */

/* ======================================================================== */
/* Rayleigh damping and element masses: Element::getDamp / getRayleighDampingForces (element/Element.cpp:182,284), */
/* Brick::formInertiaTerms (Brick.cpp:595), FourNodeQuad::getMass (FourNodeQuad.cpp:387),                           */
/* getResistingForceIncInertia of each element, getInitialStiff of each element                                    */
/* ======================================================================== */
static int ele_nd(const OrcEle* el) { return el->nen * el->ndf_e; }

/* ForceBeamColumn2d::getInitialFlexibility + Matrix::Invert (ForceBeamColumn2d.cpp getInitialStiff) */
static void beam_initial_kv(const OrcBeam* b, double* kv0) {
  double xi[ORC_MAXSEC], wt[ORC_MAXSEC], f[9];
  if (b->user_rule) { memcpy(xi, b->rxi, sizeof(double) * b->nip); memcpy(wt, b->rwt, sizeof(double) * b->nip); } else lobatto(b->nip, xi, wt);
  for (int i = 0; i < 9; i++) f[i] = 0.0;
  for (int i = 0; i < b->nip; i++) {
    double fSec[4], fb[6];
    sec_initial_flex(&b->sec[i], fSec);
    double xL = xi[i], xL1 = xL - 1.0, wtL = wt[i] * b->L;
    for (int q = 0; q < 6; q++) fb[q] = 0.0;
    for (int jj = 0; jj < 2; jj++) fb[jj + 2 * 0] += fSec[jj + 2 * 0] * wtL;
    for (int jj = 0; jj < 2; jj++) { double tmp = fSec[jj + 2 * 1] * wtL; fb[jj + 2 * 1] += xL1 * tmp; fb[jj + 2 * 2] += xL * tmp; }
    for (int jj = 0; jj < 3; jj++) f[0 + 3 * jj] += fb[0 + 2 * jj];
    for (int jj = 0; jj < 3; jj++) { double tmp = fb[1 + 2 * jj]; f[1 + 3 * jj] += xL1 * tmp; f[2 + 3 * jj] += xL * tmp; }
  }
  inv3(f, kv0);
}
/* ForceBeamColumn3d::getInitialFlexibility (ForceBeamColumn3d.cpp:2003) + Matrix::Invert */
static void beam3_initial_kv(const OrcBeam3* b, double* kv0) {
  double xi[ORC_MAXSEC], wt[ORC_MAXSEC], f[36];
  if (b->user_rule) { memcpy(xi, b->rxi, sizeof(double) * b->nip); memcpy(wt, b->rwt, sizeof(double) * b->nip); } else lobatto(b->nip, xi, wt);
  for (int i = 0; i < 36; i++) f[i] = 0.0;
  for (int i = 0; i < b->nip; i++) {
    double fSec[16], fb[24];
    sec3_initial_flex(&b->sec[i], fSec);
    double xL = xi[i], xL1 = xL - 1.0, wtL = wt[i] * b->L;
    for (int q = 0; q < 24; q++) fb[q] = 0.0;
    for (int jj = 0; jj < 4; jj++) fb[jj + 4 * 0] += fSec[jj + 4 * 0] * wtL;
    for (int jj = 0; jj < 4; jj++) { double tmp = fSec[jj + 4 * 1] * wtL; fb[jj + 4 * 1] += xL1 * tmp; fb[jj + 4 * 2] += xL * tmp; }
    for (int jj = 0; jj < 4; jj++) { double tmp = fSec[jj + 4 * 2] * wtL; fb[jj + 4 * 3] += xL1 * tmp; fb[jj + 4 * 4] += xL * tmp; }
    for (int jj = 0; jj < 4; jj++) fb[jj + 4 * 5] += fSec[jj + 4 * 3] * wtL;
    for (int jj = 0; jj < 6; jj++) f[0 + 6 * jj] += fb[0 + 4 * jj];
    for (int jj = 0; jj < 6; jj++) { double tmp = fb[1 + 4 * jj]; f[1 + 6 * jj] += xL1 * tmp; f[2 + 6 * jj] += xL * tmp; }
    for (int jj = 0; jj < 6; jj++) { double tmp = fb[2 + 4 * jj]; f[3 + 6 * jj] += xL1 * tmp; f[4 + 6 * jj] += xL * tmp; }
    for (int jj = 0; jj < 6; jj++) f[5 + 6 * jj] += fb[3 + 4 * jj];
  }
  inv6_flex(f, kv0);
}

/* Element::getTangentStiff (which = 1) / getInitialStiff (which = 2), row-major nd x nd */
static void ele_K(OrcModel* m, OrcEle* el, int which, double* K) {
  double R[24];
  if (el->kind == ORC_ELE_BRICK) brick_form(m, el, which, K, R);
  else if (el->kind == ORC_ELE_QUAD) quad_form(m, el, which, K, R);
  else if (el->kind == ORC_ELE_FBC2D) {
    if (which == 1) beam_form(el->beam, K, R);
    else { OrcBeam t = *el->beam; beam_initial_kv(el->beam, t.kv); t.pdelta = 0; beam_form(&t, K, R); }   /* getInitialGlobalStiffMatrix: no geometric part */
  } else {
    if (which == 1) beam3_form(el->beam3, K, R);
    else { OrcBeam3 t = *el->beam3; beam3_initial_kv(el->beam3, t.kv); t.pdelta = 0; beam3_form(&t, K, R); }
  }
}

/* Element::getMass: Brick::formInertiaTerms(1) (consistent), FourNodeQuad::getMass (lumped, material rho;
 * the element's own rho parameter must be 0), force beams with rho = 0 (zero matrix).
 * inertia != NULL: also the addend of formInertiaTerms to the residual (brick), row-major M. */
static void ele_M(OrcModel* m, OrcEle* el, double* M, double* inertia) {
  const int nd = ele_nd(el);
  for (int i = 0; i < nd * nd; i++) M[i] = 0.0;
  if (inertia) for (int i = 0; i < nd; i++) inertia[i] = 0.0;
  if (el->kind == ORC_ELE_BRICK) {
    double xl[3][8];
    for (int i = 0; i < 8; i++) for (int d = 0; d < 3; d++) xl[d][i] = m->crd[el->node[i] * 3 + d];
    const double one_over_root3 = 1.0 / sqrt(3.0);
    const double sg[2] = { -one_over_root3, one_over_root3 };
    double Shape[8][4][8], dvol[8];
    int count = 0;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
      double gp[3] = { sg[i], sg[j], sg[k] }, xsj;
      shp3d(gp, &xsj, Shape[count], xl);
      dvol[count] = 1.0 * xsj;
      count++;
    }
    for (int i = 0; i < 8; i++) {
      double (*shp)[8] = Shape[i];
      double momentum[3] = {0.0, 0.0, 0.0};
      for (int j = 0; j < 8; j++) {
        const double* a = m->acc + (size_t)el->node[j] * m->ndf;
        for (int p = 0; p < 3; p++) momentum[p] += a[p] * shp[3][j];      /* Vector::addVector(1.0, accel, shp) */
      }
      const double rho = el->gp[i].rho;
      for (int p = 0; p < 3; p++) momentum[p] *= rho;
      int jj = 0;
      for (int j = 0; j < 8; j++) {
        double temp = shp[3][j] * dvol[i];
        if (inertia) for (int p = 0; p < 3; p++) inertia[jj + p] += temp * momentum[p];
        temp *= rho;
        int kk = 0;
        for (int k = 0; k < 8; k++) {
          const double massJK = temp * shp[3][k];
          for (int p = 0; p < 3; p++) M[(jj + p) * 24 + kk + p] += massJK;
          kk += 3;
        }
        jj += 3;
      }
    }
  } else if (el->kind == ORC_ELE_FBC2D || el->kind == ORC_ELE_FBC3D) {
    /* ForceBeamColumn2d::getMass (ForceBeamColumn2d.cpp) / ForceBeamColumn3d::getMass: lumped, 0.5 rho L on the translations */
    const int b3 = el->kind == ORC_ELE_FBC3D;
    const double rho = el->par[b3 ? 7 : 4];
    if (rho == 0.0) return;
    const double L = b3 ? el->beam3->L : el->beam->L;
    const int ndfe = b3 ? 6 : 3, ntr = b3 ? 3 : 2;
    for (int a = 0; a < 2; a++) for (int p = 0; p < ntr; p++) { const int i = a * ndfe + p; M[i * nd + i] = 0.5 * L * rho; }
  } else if (el->kind == ORC_ELE_QUAD) {
    double rhoi[4], sum = 0.0;
    for (int i = 0; i < 4; i++) { rhoi[i] = el->gp[i].rho; sum += rhoi[i]; }
    if (sum == 0.0) return;
    for (int i = 0; i < 4; i++) {
      double shp[3][4];
      double rhodvol = quad_shape(m, el, quad_pts[i][0], quad_pts[i][1], shp);
      rhodvol *= (rhoi[i] * el->par[0] * quad_wts[i]);
      for (int alpha = 0, ia = 0; alpha < 4; alpha++, ia++) {
        const double Nrho = shp[2][alpha] * rhodvol;
        M[ia * 8 + ia] += Nrho;
        ia++;
        M[ia * 8 + ia] += Nrho;
      }
    }
  }
}

/* Element::getDamp (Element.cpp:182): alphaM M + betaK Kt + betaK0 K0 + betaKc Kc, accumulated in that order */
static void ele_damp(OrcModel* m, OrcEle* el, double* D) {
  const int nd = ele_nd(el), n2 = nd * nd;
  double T[576];
  for (int i = 0; i < n2; i++) D[i] = 0.0;
  if (m->e_alphaM != 0.0) { ele_M(m, el, T, NULL); for (int i = 0; i < n2; i++) D[i] = T[i] * m->e_alphaM; }
  if (m->betaK != 0.0) { ele_K(m, el, 1, T); for (int i = 0; i < n2; i++) D[i] += T[i] * m->betaK; }
  if (m->betaK0 != 0.0) { ele_K(m, el, 2, T); for (int i = 0; i < n2; i++) D[i] += T[i] * m->betaK0; }
  if (m->betaKc != 0.0 && el->Kc) { for (int i = 0; i < n2; i++) D[i] += el->Kc[i] * m->betaKc; }
}
static int any_rayleigh(const OrcModel* m) { return m->e_alphaM != 0.0 || m->betaK != 0.0 || m->betaK0 != 0.0 || m->betaKc != 0.0; }

/* Element::getResistingForceIncInertia as each element implements it (Brick.cpp:568, FourNodeQuad.cpp:556,
 * ForceBeamColumn2d/3d with rho = 0).  R holds getResistingForce on entry. */
static void ele_add_inertia_and_damping(OrcModel* m, OrcEle* el, double* R) {
  const int nd = ele_nd(el);
  double M[576], D[576], F[24], v[24];
  int add_damp = 0;
  if (el->kind == ORC_ELE_BRICK) {
    double inertia[24], sum = 0.0;
    for (int i = 0; i < 8; i++) sum += fabs(el->gp[i].rho);
    if (sum != 0.0) {                                      /* rho = 0: formInertiaTerms adds exact zeros */
      ele_M(m, el, M, inertia);
      for (int i = 0; i < nd; i++) R[i] += inertia[i];     /* formInertiaTerms adds into resid */
    }
    add_damp = any_rayleigh(m);
  } else if (el->kind == ORC_ELE_QUAD) {
    double sum = 0.0;
    for (int i = 0; i < 4; i++) sum += el->gp[i].rho;
    if (sum == 0.0) add_damp = (m->betaK != 0.0 || m->betaK0 != 0.0 || m->betaKc != 0.0);
    else {
      ele_M(m, el, M, NULL);
      for (int a = 0; a < 4; a++) for (int p = 0; p < 2; p++) {
        const int i = 2 * a + p;
        R[i] += M[i * 8 + i] * m->acc[(size_t)el->node[a] * m->ndf + p];
      }
      add_damp = any_rayleigh(m);
    }
  } else {
    /* ForceBeamColumn2d / 3d::getResistingForceIncInertia: m a on the translations when rho != 0, then the damping forces */
    const int b3 = el->kind == ORC_ELE_FBC3D;
    const double rho = el->par[b3 ? 7 : 4];
    if (rho != 0.0) {
      const double L = b3 ? el->beam3->L : el->beam->L;
      const double mm = 0.5 * rho * L;
      const int ndfe = b3 ? 6 : 3, ntr = b3 ? 3 : 2;
      for (int a = 0; a < 2; a++) for (int p = 0; p < ntr; p++) R[a * ndfe + p] += mm * m->acc[(size_t)el->node[a] * m->ndf + p];
      add_damp = any_rayleigh(m);
    } else add_damp = (m->betaK != 0.0 || m->betaK0 != 0.0 || m->betaKc != 0.0);
  }
  if (!add_damp) return;
  /* Element::getRayleighDampingForces: F = D v, column by column (Vector::addMatrixVector(0.0, D, v, 1.0)) */
  ele_damp(m, el, D);
  for (int a = 0; a < el->nen; a++) for (int p = 0; p < el->ndf_e; p++) v[a * el->ndf_e + p] = m->vel[(size_t)el->node[a] * m->ndf + p];
  for (int i = 0; i < nd; i++) F[i] = 0.0;
  for (int j = 0; j < nd; j++) { const double vj = v[j]; for (int i = 0; i < nd; i++) F[i] += D[i * nd + j] * vj; }
  for (int i = 0; i < nd; i++) R[i] += F[i];
}

/* `rayleigh alphaM betaK betaKinit betaKcomm` = Domain::setRayleighDampingFactors (Domain.cpp:1858): every
 * element (Element::setRayleighDampingFactors: Kc = copy of the current tangent when betaKc != 0) and every node */
int orc_set_rayleigh(void* h, double alphaM, double betaK, double betaK0, double betaKc) {
  OrcModel* m = (OrcModel*)h;
  m->e_alphaM = alphaM; m->betaK = betaK; m->betaK0 = betaK0; m->betaKc = betaKc;
  m->alphaM = alphaM;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (betaKc != 0.0) {
      if (!el->Kc) { el->Kc = (double*)malloc(sizeof(double) * 576); ele_K(m, el, 1, el->Kc); }
    } else if (el->Kc) { free(el->Kc); el->Kc = NULL; }
  }
  return 0;
}

int orc_ele_tangent(void* h, int e, double* K) {
  OrcModel* m = (OrcModel*)h; double R[24];
  if (m->ele[e].kind == ORC_ELE_FBC2D) { beam_form(m->ele[e].beam, K, R); return 6; }
  if (m->ele[e].kind == ORC_ELE_FBC3D) { beam3_form(m->ele[e].beam3, K, R); return 12; }
  if (m->ele[e].kind == ORC_ELE_BRICK) { brick_form(m, &m->ele[e], 1, K, R); return 24; }
  quad_form(m, &m->ele[e], 1, K, R); return 8;
}
int orc_ele_resid(void* h, int e, double* R) {
  OrcModel* m = (OrcModel*)h;
  if (m->ele[e].kind == ORC_ELE_FBC2D) { beam_form(m->ele[e].beam, NULL, R); return 6; }
  if (m->ele[e].kind == ORC_ELE_FBC3D) { beam3_form(m->ele[e].beam3, NULL, R); return 12; }
  if (m->ele[e].kind == ORC_ELE_BRICK) { brick_form(m, &m->ele[e], 0, NULL, R); return 24; }
  double K[64]; quad_form(m, &m->ele[e], 0, K, R); return 8;
}

/* IncrementalIntegrator::formTangent (IncrementalIntegrator.cpp:74-102):
 * zeroA; for each FE_Element: addA(getTangent, getID).  getTangent =
 * StaticIntegrator::formEleTangent (StaticIntegrator.cpp:80-97): zeroTangent; addKtToTang(1.0)
 * -> Matrix::addMatrix(K, 1.0): tang = 0*1.0 + K (adds K to a zeroed matrix). */
int orc_form_tangent(void* h, double* A) {
  OrcModel* m = (OrcModel*)h;
  memset(m->A, 0, sizeof(double) * m->nnz);
  if (m->store_kind) memset(m->Astore, 0, sizeof(double) * m->astore_size);
  /* TransientIntegrator::formTangent (TransientIntegrator.cpp:89-97): DOF_Group tangents first,
   * Newmark::formNodTangent: zeroTangent; addCtoTang(c2) (C = alphaM*M, Node::getDamp); addMtoTang(c3) */
  if (m->c2 != 0.0 || m->c3 != 0.0)
    for (int n = 0; n < m->nn; n++)
      for (int j = 0; j < m->ndf; j++) {
        int r = m->id[n * m->ndf + j];
        if (r < 0) continue;
        double t = 0.0;
        t += (m->mass[n * m->ndf + j] * m->alphaM) * m->c2;
        t += m->mass[n * m->ndf + j] * m->c3;
        if (m->store_kind) { m->Astore[store_loc(m, r, r)] += t; continue; }
        int k = soe_find(m, r, r);
        m->A[k] += t;
      }
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e]; double K[576], R[24]; int ids[32];
    int nd_e = el->nen * el->ndf_e;
    if (el->kind == ORC_ELE_BRICK) brick_form(m, el, 1, K, R);
    else if (el->kind == ORC_ELE_QUAD) quad_form(m, el, 1, K, R);
    else if (el->kind == ORC_ELE_FBC3D) beam3_form(el->beam3, K, R);
    else beam_form(el->beam, K, R);
    int n = ele_ids(m, el, ids); (void)n;
    /* FE_Element::addKtToTang(c1): theTangent->addMatrix(K, c1) on a zeroed matrix */
    for (int i = 0; i < nd_e; i++) for (int j = 0; j < nd_e; j++)
      K[i * nd_e + j] = (m->c1 == 1.0) ? 0.0 + K[i * nd_e + j] : 0.0 + K[i * nd_e + j] * m->c1;
    /* Newmark::formEleTangent (Newmark.cpp:262): addCtoTang(c2) -> += getDamp() * c2, addMtoTang(c3) -> += getMass() * c3
     * (FE_Element.cpp:279,292); both add exact zeros when the element has neither damping factors nor mass */
    if (m->c2 != 0.0 && any_rayleigh(m)) {
      double D[576]; ele_damp(m, el, D);
      for (int i = 0; i < nd_e * nd_e; i++) K[i] += D[i] * m->c2;
    }
    if (m->c3 != 0.0) {      /* (a force beam without -mass: a zero matrix) */
      double M[576]; ele_M(m, el, M, NULL);
      for (int i = 0; i < nd_e * nd_e; i++) K[i] += M[i] * m->c3;
    }
    if (m->store_kind) {
      /* BandGenLinSOE::addA / ProfileSPDLinSOE::addA: for (i) col = id(i); for (j) row = id(j); *APtr += m(j,i) */
      for (int i = 0; i < nd_e; i++) { int col = ids[i]; if (col < 0) continue;
        for (int j = 0; j < nd_e; j++) { int row = ids[j]; if (row < 0) continue;
          long long k = store_loc(m, row, col); if (k >= 0) m->Astore[k] += K[j * nd_e + i]; } }
    } else if (m->soe_kind == 1) {
      for (int i = 0; i < nd_e; i++) { int row = ids[i]; if (row < 0) continue;
        for (int j = 0; j < nd_e; j++) { int col = ids[j]; if (col < 0) continue;
          int k = soe_find(m, row, col); if (k >= 0) m->A[k] += K[i * nd_e + j]; } }
    } else {
      for (int i = 0; i < nd_e; i++) { int col = ids[i]; if (col < 0) continue;
        for (int j = 0; j < nd_e; j++) { int row = ids[j]; if (row < 0) continue;
          int k = soe_find(m, col, row); if (k >= 0) m->A[k] += K[j * nd_e + i]; } }
    }
  }
  if (A) { if (m->store_kind) memcpy(A, m->Astore, sizeof(double) * m->astore_size); else memcpy(A, m->A, sizeof(double) * m->nnz); }
  return 0;
}

/* IncrementalIntegrator::formUnbalance: zeroB; formElementResidual
 * (IncrementalIntegrator.cpp:221-238) then formNodalUnbalance (202-218).
 * FE_Element::getResidual -> StaticIntegrator::formEleResidual
 * (StaticIntegrator.cpp:100-108): zeroResidual; addRtoResidual(1.0) ->
 * theResidual->addVector(1.0, R, -1.0) (FE_Element.cpp:401-414).
 * DOF_Group::getUnbalance -> Node::getUnbalancedLoad = lambda * load
 * (NodalLoad::applyLoad -> Node::addUnbalancedLoad: unbal += load*fact). */
int orc_form_unbalance(void* h, double* B) {
  OrcModel* m = (OrcModel*)h;
  memset(m->B, 0, sizeof(double) * (m->neq > 0 ? m->neq : 1));
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e]; double K[64], R[24]; int ids[32];
    int nd_e = el->nen * el->ndf_e;
    if (el->kind == ORC_ELE_BRICK) brick_form(m, el, 0, NULL, R);
    else if (el->kind == ORC_ELE_QUAD) quad_form(m, el, 0, K, R);
    else if (el->kind == ORC_ELE_FBC3D) beam3_form(el->beam3, NULL, R);
    else beam_form(el->beam, NULL, R);
    /* TransientIntegrator::formEleResidual -> addRIncInertiaToResidual -> getResistingForceIncInertia; with
     * zero velocities / accelerations (static analysis) the extra terms are exact zeros */
    ele_add_inertia_and_damping(m, el, R);
    ele_ids(m, el, ids);
    for (int i = 0; i < nd_e; i++) {
      double res = 0.0 * 1.0 + R[i] * -1.0;   /* Vector::addVector(1.0, R, -1.0) on a zeroed residual */
      if (ids[i] >= 0) m->B[ids[i]] += res;
    }
  }
  for (int n = 0; n < m->nn; n++)
    for (int j = 0; j < m->ndf; j++) {
      int pos = m->id[n * m->ndf + j];
      if (pos < 0) continue;
      /* Node::getUnbalancedLoadIncInertia (Node.cpp): P - M a - alphaM M v (static: c2 = c3 = 0 and v = a = 0) */
      /* Domain::applyLoad: pattern after pattern, NodalLoad::applyLoad -> Node::addUnbalancedLoad(load, factor);
       * a constant pattern keeps the factor it had at loadConst (LoadPattern.cpp applyLoad) */
      double ub = m->cload ? (0.0 + m->cload[n * m->ndf + j]) + m->load[n * m->ndf + j] * m->lambda
                           : 0.0 + m->load[n * m->ndf + j] * m->lambda;
      double ms = m->mass[n * m->ndf + j];
      if (ms != 0.0) { ub -= ms * m->acc[n * m->ndf + j]; if (m->alphaM != 0.0) ub += ms * m->vel[n * m->ndf + j] * -m->alphaM; }
      m->B[pos] += ub;
    }
  if (B) memcpy(B, m->B, sizeof(double) * m->neq);
  return 0;
}

/* Domain::commit -> Element::commitState -> material commitState; nodes commit trial */
int orc_commit(void* h) {
  OrcModel* m = (OrcModel*)h;
  m->lambda_c = m->lambda;                                    /* Domain::commit: committedTime = currentTime */
  memcpy(m->commit_disp, m->trial, sizeof(double) * m->nn * m->ndf);
  memset(m->incr, 0, sizeof(double) * m->nn * m->ndf);        /* Node::commitState */
  memcpy(m->velc, m->vel, sizeof(double) * m->nn * m->ndf); memcpy(m->accc, m->acc, sizeof(double) * m->nn * m->ndf);
  for (int e = 0; e < m->ne; e++) {
    if (m->ele[e].Kc) ele_K(m, &m->ele[e], 1, m->ele[e].Kc);   /* Element::commitState: *Kc = getTangentStiff(), before the materials commit */
    if (m->ele[e].kind == ORC_ELE_FBC2D) beam_commit(m->ele[e].beam);
    else if (m->ele[e].kind == ORC_ELE_FBC3D) beam3_commit(m->ele[e].beam3);
    else for (int g = 0; g < m->ele[e].nip; g++) gp_commit(&m->ele[e].gp[g]);
  }
  return 0;
}
/* Domain::revertToStart (Domain.cpp:1951, the `reset` command): nodes and elements back to their initial state
 * (Node::revertToStart; Brick / FourNodeQuad -> NDMaterial::revertToStart: J2Plasticity.cpp:532 zero(), J2PlaneStress also
 * commitEps22 = 0; ForceBeamColumn2d.cpp:344 / 3d: sections, fs, vs, Ssr, Se, kv zero, initialFlag = 0), time and
 * load factor 0, applyLoad(0), update().  Element::Kc (Rayleigh betaKc) is NOT reset by the reference. */
int orc_revert_to_start(void* h) {
  OrcModel* m = (OrcModel*)h;
  const size_t nb = sizeof(double) * m->nn * m->ndf;
  memset(m->trial, 0, nb); memset(m->commit_disp, 0, nb); memset(m->incr, 0, nb);
  memset(m->vel, 0, nb); memset(m->velc, 0, nb); memset(m->acc, 0, nb); memset(m->accc, 0, nb);
  m->lambda = m->lambda_c = 0.0;
  for (int e = 0; e < m->ne; e++) {
    OrcEle* el = &m->ele[e];
    if (el->kind == ORC_ELE_FBC2D) {
      for (int i = 0; i < el->beam->nip; i++) free(el->beam->sec[i].mat);
      OrcBeam keep = *el->beam;      /* the element loads are not part of the state: they stay as addLoad left them */
      free(el->beam);
      el->beam = beam2_build(m, el, el->mat, el->par);
      el->beam->has_load = keep.has_load; el->beam->numEleLoads = keep.numEleLoads; el->beam->loadFactor = keep.loadFactor;
      memcpy(el->beam->w, keep.w, sizeof keep.w);
      el->beam->has_point = keep.has_point; memcpy(el->beam->pt, keep.pt, sizeof keep.pt);
      el->beam->has_partial = keep.has_partial; memcpy(el->beam->pq, keep.pq, sizeof keep.pq);
      el->beam->user_rule = keep.user_rule; memcpy(el->beam->rxi, keep.rxi, sizeof keep.rxi); memcpy(el->beam->rwt, keep.rwt, sizeof keep.rwt);
    } else if (el->kind == ORC_ELE_FBC3D) {
      for (int i = 0; i < el->beam3->nip; i++) free(el->beam3->sec[i].mat);
      const int hl = el->beam3->has_load, nl = el->beam3->numEleLoads; const double lf = el->beam3->loadFactor;
      double wk[3]; memcpy(wk, el->beam3->w, sizeof wk);
      const int hp = el->beam3->has_point; double pk[4]; memcpy(pk, el->beam3->pt, sizeof pk);
      const int hq = el->beam3->has_partial; double qk[8]; memcpy(qk, el->beam3->pq, sizeof qk);
      const int ur = el->beam3->user_rule; double rx[ORC_MAXSEC], rw[ORC_MAXSEC]; memcpy(rx, el->beam3->rxi, sizeof rx); memcpy(rw, el->beam3->rwt, sizeof rw);
      free(el->beam3);
      el->beam3 = beam3_build(m, el, el->mat, el->par);
      el->beam3->has_load = hl; el->beam3->numEleLoads = nl; el->beam3->loadFactor = lf; memcpy(el->beam3->w, wk, sizeof wk);
      el->beam3->has_point = hp; memcpy(el->beam3->pt, pk, sizeof pk);
      el->beam3->has_partial = hq; memcpy(el->beam3->pq, qk, sizeof qk);
      el->beam3->user_rule = ur; memcpy(el->beam3->rxi, rx, sizeof rx); memcpy(el->beam3->rwt, rw, sizeof rw);
    } else {
      for (int g = 0; g < el->nip; g++) {
        OrcGP* gp = &el->gp[g];
        if (gp->kind == ORC_MAT_J2) {   /* J2Plasticity::zero: history, stress, strain (the tangent stays until the update) */
          j2_zero(&gp->u.j2);
          memset(gp->u.j2.stress, 0, sizeof gp->u.j2.stress); memset(gp->u.j2.strain, 0, sizeof gp->u.j2.strain);
          gp->u.j2.commitEps22 = 0.0;
        } else { memset(gp->u.el.epsilon, 0, sizeof gp->u.el.epsilon); memset(gp->u.el.Cepsilon, 0, sizeof gp->u.el.Cepsilon); }
      }
    }
  }
  orc_apply_load(h, 0.0);       /* Domain::revertToStart: applyLoad(currentTime = 0), then update() */
  int rc = 0;
  for (int e = 0; e < m->ne; e++) rc |= ele_update(m, &m->ele[e]);
  return rc;
}

int orc_revert(void* h) {
  OrcModel* m = (OrcModel*)h;
  /* Domain::revertToLastCommit (Domain.cpp:1925): nodes, elements, currentTime = committedTime + applyLoad, then update() */
  orc_apply_load(h, m->lambda_c);
  memcpy(m->trial, m->commit_disp, sizeof(double) * m->nn * m->ndf);
  memset(m->incr, 0, sizeof(double) * m->nn * m->ndf);
  memcpy(m->vel, m->velc, sizeof(double) * m->nn * m->ndf); memcpy(m->acc, m->accc, sizeof(double) * m->nn * m->ndf);
  for (int e = 0; e < m->ne; e++) {
    if (m->ele[e].kind == ORC_ELE_FBC2D) beam_revert(m->ele[e].beam);
    else if (m->ele[e].kind == ORC_ELE_FBC3D) beam3_revert(m->ele[e].beam3);
    else for (int g = 0; g < m->ele[e].nip; g++) gp_revert(&m->ele[e].gp[g]);
  }
  int rc = 0;
  for (int e = 0; e < m->ne; e++) rc |= ele_update(m, &m->ele[e]);
  return rc;
}

/* per-GP peek for kernel-level parity: stress (order) and tangent (order^2) of FE element e, point g */
int orc_gp_response(void* h, int e, int g, double* stress, double* tangent) {
  OrcModel* m = (OrcModel*)h; OrcGP* gp = &m->ele[e].gp[g];
  gp_get_stress(gp, stress); gp_get_tangent(gp, tangent);
  return gp->type == ORC_ND_3D ? 6 : 3;
}

void orc_model_free(void* h) {
  OrcModel* m = (OrcModel*)h; if (!m) return;
  free(m->node_tag); free(m->crd); free(m->trial); free(m->commit_disp); free(m->load); free(m->fixed);
  free(m->mat_tag); free(m->mat_kind); free(m->mat_par); free(m->ele); free(m->id); free(m->ptr); free(m->idx);
  free(m->A); free(m->B); free(m->iDiagLoc); free(m->Astore); free(m->node_ndf); free(m);
}
