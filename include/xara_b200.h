/* xara_b200 -- B200-native element state determination + global assembly.
 *
 * C-ABI drop-in boundary for the reference's (peer-open-source/xara, OpenSeesRT)
 * hot path:  Domain::update -> Element::update -> NDMaterial::setTrialStrain,
 * IncrementalIntegrator::formTangent / formUnbalance -> FE_Element::getTangent /
 * getResidual -> LinearSOE::addA / addB.
 *
 * Plain pointers and sizes only.  Every entry point names the reference
 * interface it stands in for (path under /root/reference/SRC : line).
 * INTEGRATION.md shows the C++ glue a maintainer adds on the reference side.
 *
 * There is NO CPU fallback: every xb_* call that needs the device fails with
 * XB_ERR_CUDA when no sm_100a device is available.
 */
#ifndef XARA_B200_H
#define XARA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xb_model xb_model;

/* return codes */
enum {
  XB_OK = 0,
  XB_ERR_ARG = -1,        /* bad argument (unknown tag, wrong size, order ...)   */
  XB_ERR_STATE = -2,      /* call made in the wrong phase (e.g. before xb_setup) */
  XB_ERR_CUDA = -3,       /* CUDA runtime error or no device                      */
  XB_ERR_MATERIAL = -4,   /* constitutive update failed (J2 return map > 25 its,  */
                          /*   J2Plasticity.cpp:296-301 returns -1)               */
  XB_ERR_UNSUPPORTED = -5 /* model feature outside the device path               */
};

/* nDMaterial kinds (runtime/commands/modeling/nDMaterial.cpp, material/plastic.cpp:927) */
enum {
  XB_MAT_ELASTIC_ISOTROPIC = 0, /* par = E, nu, rho          (ElasticIsotropicMaterial.h)   */
  XB_MAT_J2PLASTICITY = 1       /* par = K, G, sig0, sigInf, delta, H, eta (J2Plasticity.h:47) */
};

/* uniaxialMaterial kinds (fibres of a fibre section) */
enum {
  XB_UNI_STEEL02 = 0,   /* material/uniaxial/steel/Steel02.h:47: Fy,E0,b,R0,cR1,cR2,a1,a2,a3,a4[,sigInit] */
  XB_UNI_CONCRETE02 = 1,/* material/uniaxial/concrete/Concrete02.cpp:93: fc,epsc0,fcu,epscu,rat,ft,Ets     */
  XB_UNI_STEEL01 = 2,   /* material/uniaxial/steel/Steel01.cpp:40: fy,E0,b,a1,a2,a3,a4 (the command's defaults 0,55,0,55) */
  XB_UNI_ELASTIC = 3,   /* material/uniaxial/ElasticMaterial.cpp:96: E[,eta,Eneg]; eta must be 0 (no strain rate on this path) */
  XB_UNI_CONCRETE01 = 4,/* material/uniaxial/concrete/Concrete01.cpp:89: fpc,epsc0,fpcu,epscu (Kent-Scott-Park, no tension) */
  XB_UNI_ELASTICPP = 5  /* material/uniaxial/ElasticPPMaterial.cpp:88: E,epsyP,epsyN,eps0 (elastic - perfectly plastic; the
                           plastic strain moves at commitState, :190-224) */
};

/* element kinds */
enum {
  XB_ELE_STDBRICK = 0,    /* element/Brick/Brick.cpp, 8 nodes x 3 dof, 2x2x2 Gauss; par = b1,b2,b3 */
  XB_ELE_FOURNODEQUAD = 1,/* element/Plane/FourNodeQuad.cpp, 4 nodes x 2 dof, 2x2 Gauss;
                             par = thickness, type (0 PlaneStrain, 1 PlaneStress; with J2Plasticity the copy is
                             J2PlaneStrain / J2PlaneStress: one type per xb_add_elements call), surface pressure
                             (setPressureLoadAtNodes, FourNodeQuad.cpp:1206), rho (unused: the material's), b1, b2 */
  XB_ELE_FORCEBEAMCOLUMN2D = 2,/* element/Frame/Other/Force/ForceBeamColumn2d.cpp, 2 nodes x 3 dof; mat_tags name the
                             section (xb_add_fiber_section or xb_add_section_aggregator);
                             par = nIP, maxIters, tol [, geomTransf: 0 Linear | 1 PDelta (coordTransformation/
                             PDeltaCrdTransf2d.cpp) | 2 Corotational (CorotCrdTransf2d.cpp; no joint offsets, no
                             stiffness-proportional Rayleigh terms: XB_ERR_UNSUPPORTED) [, rho: `-mass`, mass per unit length, lumped [, dXi, dYi, dXj,
                             dYj: `-jntOffset`, rigid end zones]]] -- rows of 3, 4, 5 or 9 values; one
                             section/nIP/maxIters/tol/geomTransf per call.  Lobatto
                             integration unless xb_set_beam_integration hands over another rule's points          */
  XB_ELE_FORCEBEAMCOLUMN3D = 3 /* `element forceBeamColumn` in a 3D model (runtime/commands/modeling/element/
                             frames.cpp:333) = element/Frame/Other/Force/ForceBeamColumn3d.cpp, 2 nodes x 6 dof,
                             Lobatto integration unless xb_set_beam_integration hands over another rule; mat_tags
                             name a section added with xb_add_fiber_section3d;
                             par = nIP, maxIters, tol, vecxz[3] [, geomTransf: 0 Linear | 1 PDelta (LinearCrdTransf3d /
                             PDeltaCrdTransf3d) [, rho [, dXi, dYi, dZi, dXj, dYj, dZj: `-jntOffset`]]] -- rows of
                             6, 7, 8 or 14 values (one section/nIP/maxIters/tol/geomTransf per call)  */
};

/* DOF numberers (analysis/numberer) */
enum {
  XB_NUMBERER_PLAIN = 0, /* PlainNumberer::numberDOF, PlainNumberer.cpp:73            */
  XB_NUMBERER_RCM = 1    /* DOF_Numberer::numberDOF + RCM(false), DOF_Numberer.cpp:92, RCM.cpp:66 */
};

/* LinearSOE storage whose addA semantics the scatter map reproduces */
enum {
  XB_SOE_SPARSE_GEN_COL = 0, /* SparseGenColLinSOE (colStartA,rowA), SparseGenColLinSOE.cpp:161,264 */
  XB_SOE_SPARSE_GEN_ROW = 1, /* SparseGenRowLinSOE (rowStartA,colA), SparseGenRowLinSOE.cpp:127,224 */
  /* `system BandGeneral`: BandGenLinSOE (bandGEN/BandGenLinSOE.cpp:116 setSize, :208 addA).  A is the LAPACK band
   * array, ldA = 2 numSubD + numSuperD + 1 doubles per column: entry (row, col) at col ldA + numSubD + numSuperD + row - col */
  XB_SOE_BAND_GEN = 2,
  /* `system ProfileSPD`: ProfileSPDLinSOE (profileSPD/ProfileSPDLinSOE.cpp:115 setSize, :214 addA).  Upper profile by
   * columns, iDiagLoc[col] (1-based) = location of the diagonal: entry (row <= col, col) at iDiagLoc[col] - 1 + row - col */
  XB_SOE_PROFILE_SPD = 3,
  /* `system Umfpack`: UmfpackGenLinSOE (umfGEN/UmfpackGenLinSOE.cpp:70 setSize, :142 addA): Ap, Ai, Ax -- the same
   * sorted compressed columns as SparseGenCol */
  XB_SOE_UMFPACK_GEN = 4
};

const char* xb_version(void);
/* message for the last failing call on this thread */
const char* xb_last_error(void);
/* number of usable CUDA devices (0 on a CPU-only box; never an error) */
int xb_device_count(void);

/* ---- model definition: what the reference's model commands put in the Domain ---- */

/* Domain::Domain (domain/domain/Domain.cpp:84); ndm, ndf as in `model basic -ndm -ndf` */
xb_model* xb_model_create(int ndm, int ndf);
void xb_model_destroy(xb_model*);

/* Domain::addNode (Domain.cpp:576) for n nodes; crd is [n][ndm]; tags unique, any order */
int xb_add_nodes(xb_model*, int n, const int* tags, const double* crd);
/* Domain::addSP_Constraint (Domain.cpp:636) -- homogeneous `fix`; dof is 0-based */
int xb_add_sp(xb_model*, int n, const int* node_tags, const int* dofs);
/* Nodes created under another `model -ndf` than the model's: they carry ndf (< the model's) dofs -- a FourNodeQuad's
 * 2-dof nodes (FourNodeQuad.cpp:133-139 insists on them) beside the 3-dof nodes of a frame, tied with `equalDOF`.  The
 * dofs such a node does not have get no equation (its DOF_Group::myID is simply shorter in the reference); every
 * [nn][ndf] array of this interface keeps the model's ndf as its stride.  An element must be connected to nodes that
 * carry exactly its dofs per node (xb_setup returns XB_ERR_ARG otherwise). */
int xb_set_node_ndf(xb_model*, int n, const int* node_tags, int ndf);
/* Domain::addMP_Constraint (Domain.cpp:739) for `equalDOF rNode cNode dofs...`
 * (runtime/commands/domain/constraint.cpp): an MP_Constraint with an identity constraint matrix, the only kind
 * PlainHandler accepts (PlainHandler.cpp:129-176).  The n (0-based) dofs of the constrained node take the
 * equation numbers of the same dofs of the retained node (PlainNumberer.cpp:111-142, DOF_Numberer.cpp:151-190).
 * On a partitioned model the nodes of a tie group are owned by one rank (the lowest owner in the group).
 * Chains (a retained dof that is itself constrained) return XB_ERR_UNSUPPORTED at set-up. */
int xb_add_equal_dof(xb_model*, int retained_node_tag, int constrained_node_tag, int n, const int* dofs);
/* OPS nDMaterial command; par has npar doubles in the order listed at the kind */
int xb_add_nd_material(xb_model*, int tag, int kind, const double* par, int npar);
/* uniaxialMaterial Steel02 | Concrete02 | Steel01 | Elastic | Concrete01 (runtime/commands/modeling/uniaxial.cpp) */
int xb_add_uniaxial_material(xb_model*, int tag, int kind, const double* par, int npar);
/* section Fiber -> FiberSection2d (material/section/FiberSection2d.cpp:99 addFiber): nf fibres
 * (y, A, uniaxial material tag) in the order given; the centroid is computed as the command does */
int xb_add_fiber_section(xb_model*, int tag, int nf, const double* y, const double* A, const int* mat_tags);
/* section Aggregator tag mat1 P mat2 Mz -> SectionAggregator without a base section (material/section/
 * SectionAggregator.cpp:119, :316 setTrialSectionDeformation, :419 getSectionFlexibility): n uniaxial materials, one
 * per section response; codes as in SectionForceDeformation.h (2 = P, 1 = Mz).  A 2D forceBeamColumn takes the pair
 * (P, Mz) in that order -- BASELINE configs[0]'s section; anything else returns XB_ERR_UNSUPPORTED */
int xb_add_section_aggregator(xb_model*, int tag, int n, const int* mat_tags, const int* codes);
/* section Fiber tag -GJ gj in a 3D model -> FiberSection3d (material/section/FiberSection3d.cpp:294 addFiber,
 * runtime/commands/modeling/section.cpp:497): fibres (y, z, A, uniaxial tag), response P, Mz, My and an
 * elastic torsion GJ */
int xb_add_fiber_section3d(xb_model*, int tag, int nf, const double* y, const double* z, const double* A,
                           const int* mat_tags, double GJ);
/* Domain::addElement (Domain.cpp:442) for n elements of one kind; conn is [n][nen] node
 * tags, mat_tags [n], par [n][par_stride] (par_stride >= the kind's parameter count).
 * All materials referenced by one call must be of one nDMaterial kind. */
int xb_add_elements(xb_model*, int kind, int n, const int* tags, const int* conn,
                    const int* mat_tags, const double* par, int par_stride);
/* LoadPattern 1 with a Linear series + NodalLoad (domain/node/NodalLoad.cpp:97):
 * values is [n][ndf]; loads on one node accumulate */
int xb_add_nodal_loads(xb_model*, int n, const int* node_tags, const double* values);
/* `eleLoad -ele tags -type -beamUniform wy [wz] wa` on forceBeamColumn elements, in the same Linear pattern as the nodal
 * loads (Beam2dUniformLoad / Beam3dUniformLoad -> ForceBeamColumn2d/3d::addLoad, ForceBeamColumn2d.cpp:1005): w = [n][3]
 * = wy, wz, wa (2D: wz ignored).  The element then carries the section forces sp inside its iteration
 * (computeSectionForces, :1034) and the fixed-end reactions p0 in its resisting force (computeReactions, :407), both
 * scaled by the load factor of xb_apply_load.  One uniform load per element; point and partial loads: XB_ERR_UNSUPPORTED
 * at the binding. */
int xb_add_beam_uniform_loads(xb_model*, int n, const int* ele_tags, const double* w);
/* `element forceBeamColumn ... -integration Legendre | Radau | NewtonCotes | Trapezoidal | UserDefined ...` (the
 * BeamIntegration classes under quadrature/Frame/): the nip section locations and weights, as fractions of the element
 * length, exactly as the element's BeamIntegration object returns them (getSectionLocations / getSectionWeights,
 * ForceBeamColumn2d.cpp:598-602) -- xi, wt are [n][nip].  Without this call an element integrates with the Lobatto
 * tables (quadrature/Frame/LobattoBeamIntegration.cpp).  All elements of one xb_add_elements call or none. */
int xb_set_beam_integration(xb_model*, int n, const int* ele_tags, int nip, const double* xi, const double* wt);
/* `eleLoad -ele tags -type -beamPoint Py [Pz] xL [N]` of the Linear pattern (Beam2dPointLoad / Beam3dPointLoad ->
 * ForceBeamColumn2d.cpp:442-455, 1138-1181; ForceBeamColumn3d.cpp:457-475, 1314-1373): p is [n][4] = Py, Pz (3D), N, xL = a/L.
 * One point load per element (beside its uniform loads, whose intensities add up); a load with xL outside [0, 1] is ignored, as the
 * element does. */
int xb_add_beam_point_loads(xb_model*, int n, const int* ele_tags, const double* p);
/* `eleLoad -beamUniform` over part of an element, trapezoidal between a = aOverL L and b = bOverL L
 * (Beam2dPartialUniformLoad -> ForceBeamColumn2d.cpp:426-443 reactions, :1073-1137 section forces;
 * Beam3dPartialUniformLoad -> ForceBeamColumn3d.cpp:432-456, 1224-1313): p is [n][8] = wy_a, wy_b, wAxial_a, wAxial_b,
 * aOverL, bOverL, wz_a, wz_b (the last two: 3D elements only).  One such load per element, beside its uniform and point
 * loads; 0 <= aOverL < bOverL <= 1. */
int xb_add_beam_partial_loads(xb_model*, int n, const int* ele_tags, const double* p);

/* `mass` command: Node::setMass with a diagonal matrix (domain/node/Node.h:127); mass is [n][ndf].
 * Element masses come from the nDMaterial density (J2Plasticity par[7], ElasticIsotropic par[2]): stdBrick forms
 * the consistent mass (Brick::formInertiaTerms, Brick.cpp:595), FourNodeQuad the lumped one (FourNodeQuad.cpp:387,
 * the element's own rho parameter must stay 0); force beam-columns carry no element mass (rho = 0). */
int xb_set_nodal_mass(xb_model*, int n, const int* node_tags, const double* mass);
/* `rayleigh alphaM 0 0 0` on the nodes (Node::setRayleighDampingFactor): C_node = alphaM * M_node */
int xb_set_rayleigh_alpha_m(xb_model*, double alphaM);
/* `rayleigh alphaM betaK betaKinit betaKcomm` = Domain::setRayleighDampingFactors (domain/domain/Domain.cpp:1858):
 * every element (Element::setRayleighDampingFactors, element/Element.cpp:110; Kc is set to the current tangent
 * when betaKcomm != 0 and refreshed at every commit) and every node.  With transient factors (c1,c2,c3) the
 * element tangent becomes c1 Kt + c2 (alphaM M + betaK Kt + betaK0 K0 + betaKc Kc) + c3 M and the element
 * residual getResistingForceIncInertia = R + M a + C v.  Any time before or after xb_device_init. */
int xb_set_rayleigh(xb_model*, double alphaM, double betaK, double betaK0, double betaKc);

/* ---- analysis set-up: BasicAnalysisBuilder::domainChanged (runtime/runtime/
 * BasicAnalysisBuilder.cpp:225): PlainHandler::handle (analysis/handler/PlainHandler.cpp:60),
 * numberDOF, AnalysisModel::getDOFGraph (analysis/model/AnalysisModel.cpp:286),
 * LinearSOE::setSize.  Host-side integer work; needs no device.  Returns numEqn >= 0. */
int xb_setup(xb_model*, int numberer, int soe_kind);

/* Multi-GPU: the same set-up for rank `rank` of `nparts` (one process per GPU).  Every rank
 * is given the WHOLE model and computes the one global numbering; elements are then split
 * (part = NULL: built-in recursive coordinate bisection; else part[e] = rank of the e-th
 * element in FE_Element order, e.g. from METIS as domain/partitioner/DomainPartitioner.cpp
 * does), a node's equations are owned by the lowest rank holding one of its elements, and the
 * model is cut down to this rank's elements, their nodes, and the COMPLETE rows of its owned
 * equations (global column numbers).  Rows of element matrices / residuals computed here for
 * nodes owned elsewhere travel in xb_exchange; the owner adds all contributions in global
 * FE_Element order, so the assembled rows equal the single-GPU ones bit for bit. */
int xb_setup_partitioned(xb_model*, int numberer, int soe_kind, int nparts, int rank, const int* part);
/* owned equations of this rank (== xb_num_eqn when unpartitioned) and their global numbers */
int xb_num_rows(const xb_model*);
int xb_get_row_eqns(const xb_model*, int* eqns);
/* partition of every element, global FE_Element order [xb_num_elements of the whole model] */
int xb_get_partition(const xb_model*, int* part);
int xb_num_peers(const xb_model*);
/* counts = {send_k, recv_k, send_r, recv_r (doubles), chunks_out, chunks_in} */
int xb_get_peer(const xb_model*, int i, int* rank, long long* counts);

int xb_num_nodes(const xb_model*);
long long xb_num_elements(const xb_model*);
long long xb_num_gauss_points(const xb_model*);
int xb_num_eqn(const xb_model*);
/* number of pattern entries of the owned rows: the length of idx in xb_get_pattern (and of A for the compressed SOEs) */
long long xb_nnz(const xb_model*);
/* length of the SOE's A array, i.e. of the buffer xb_form_tangent / xb_assemble_tangent fill: xb_nnz for the compressed
 * SOEs, size * (2 numSubD + numSuperD + 1) for BandGeneral, profileSize for ProfileSPD */
long long xb_a_size(const xb_model*);
/* BandGenLinSOE::setSize: numSubD, numSuperD (XB_SOE_BAND_GEN only) */
int xb_get_band(const xb_model*, int* numSubD, int* numSuperD);
/* ProfileSPDLinSOE::setSize: iDiagLoc[neq], 1-based as in the reference (XB_SOE_PROFILE_SPD only) */
int xb_get_profile(const xb_model*, int* iDiagLoc);
/* node tags in Domain iteration order (ascending, MapOfTaggedObjects); every [nn][..]
 * array below uses this order */
int xb_get_node_tags(const xb_model*, int* tags);
/* DOF_Group::getID for every (local) node: GLOBAL ids [nn][ndf]; -1 = constrained (PlainHandler.cpp:117) */
int xb_get_ids(const xb_model*, int* ids);
/* element tags in FE_Element order (ascending element tag, PlainHandler.cpp:233) */
int xb_get_element_tags(const xb_model*, int* tags);
/* colStartA/rowA (CSC) or rowStartA/colA (CSR): ptr [neq+1] (64-bit), idx [nnz].  BandGeneral / ProfileSPD / Umfpack
 * models return the compressed-column pattern of the DOF graph (Umfpack: its Ap / Ai) */
int xb_get_pattern(const xb_model*, long long* ptr, int* idx);
/* the addA location of every entry of FE elements [e0,e1): map[(e-e0)*nd*nd + i*nd + j]
 * = index into A (of the SOE kind given to xb_setup) that receives element-matrix entry (i,j), -1 when dropped
 * (constrained dofs; ProfileSPD: the lower triangle) */
int xb_get_scatter_map(const xb_model*, long long e0, long long e1, long long* map);

/* ---- device phase ---- */

/* allocate HBM, upload the model, bind to a CUDA stream (cudaStream_t as void*; NULL =
 * a stream the library creates).  Must follow xb_setup. */
int xb_device_init(xb_model*, int device, void* cuda_stream);

/* Node::setTrialDisp for all nodes, u [nn][ndf] in host memory (H2D inside) */
int xb_set_trial_disp(xb_model*, const double* u);
/* AnalysisModel::incrDisp (analysis/model/AnalysisModel.cpp): trial += dU[id], dU [neq] host
 * (the GLOBAL increment, also on a partitioned model) */
int xb_incr_trial_disp(xb_model*, const double* dU);
int xb_get_trial_disp(xb_model*, double* u);
/* ---- transient analysis (analysis/integrator/Dynamic/Newmark.cpp, TransientIntegrator.cpp) ----
 * c1,c2,c3 of Newmark::newStep (:117-140): formTangent gives c1*K + c2*C + c3*M with the
 * DOF_Group (nodal mass) terms added first, formUnbalance gives P - M a - C v - R.  (1,0,0) = static. */
int xb_set_transient_factors(xb_model*, double c1, double c2, double c3);
/* Newmark::newStep predictor, displacement unknown (:150-160): V = a1*V + a2*A, A = a4*A + a3*V_old */
int xb_newmark_predict(xb_model*, double a1, double a2, double a3, double a4);
/* Newmark::update (:411) + AnalysisModel::setResponse: U += cu*dU[id], V += cv*dU[id], A += ca*dU[id];
 * dU is the GLOBAL increment [neq] in host memory */
int xb_incr_trial_response(xb_model*, const double* dU, double cu, double cv, double ca);
/* AnalysisModel::setVel / setAccel and their read-back, [nn][ndf] host arrays */
int xb_set_trial_vel_accel(xb_model*, const double* v, const double* a);
int xb_get_trial_vel_accel(xb_model*, double* v, double* a);
/* Domain::update -> Element::update -> NDMaterial::setTrialStrain for every Gauss point.  The call is asynchronous:
 * a state determination that fails on the device (J2 return map or force-beam element iteration not converging, where
 * Domain::update would return < 0) raises a flag that the NEXT call with a host destination (xb_form_tangent /
 * xb_form_unbalance with a non-null buffer) or xb_synchronize returns as XB_ERR_STATE.  A caller that needs the
 * reference's immediate failure semantics calls xb_synchronize right after xb_update. */
int xb_update(xb_model*);
/* AnalysisModel::applyLoadDomain(lambda) for the Linear-series pattern: Domain::applyLoad, then the constraint handler's
 * applyLoad -- nothing under `constraints Plain`; under "constraints_transformation" (xb_set_option) the second
 * Element::update of the elements with a constrained node */
int xb_apply_load(xb_model*, double lambda);
/* the load factor alone (no handler action): for an integrator that re-reads the domain time where the reference does not
 * call applyLoadDomain, e.g. before formUnbalance */
int xb_set_load_factor(xb_model*, double lambda);
/* `loadConst` (Domain::setLoadConstant, domain/domain/Domain.cpp:1814; LoadPattern::setLoadConstant, domain/pattern/
 * LoadPattern.cpp): the nodal loads applied so far stay at the current load factor; the reference load vector is
 * emptied for the next pattern; the beam element loads (xb_add_beam_uniform_loads / xb_add_beam_point_loads / xb_add_beam_partial_loads, which all
 * belong to the patterns defined before the set-up) keep that factor too -- the gravity-then-pushover sequence of an RC
 * frame.  The caller sets the new domain time with xb_apply_load (`loadConst -time 0.0`). */
int xb_load_const(xb_model*);
/* `pattern Plain tag Linear { load node values... }` defined after the set-up (the pushover pattern that follows
 * loadConst): values [n][ndf] are added to the reference loads of the nodes */
int xb_set_nodal_loads(xb_model*, int n, const int* node_tags, const double* values);
/* IncrementalIntegrator::formTangent(CURRENT_TANGENT), analysis/integrator/
 * IncrementalIntegrator.cpp:74.  A (nnz doubles, host) may be NULL to keep A resident. */
int xb_form_tangent(xb_model*, double* A);
/* IncrementalIntegrator::formUnbalance (formElementResidual :221 + formNodalUnbalance :202).
 * B (neq doubles, host) may be NULL. */
int xb_form_unbalance(xb_model*, double* B);
/* the two halves of formTangent / formUnbalance, separately callable (profiling, overlap):
 * FE_Element::getTangent / getResidual for every element into device-resident element
 * matrices / vectors (analysis/fe_ele/FE_Element.cpp:235,370), then LinearSOE::addA / addB
 * as a fixed-order gather.  xb_form_tangent == elements + assemble. */
int xb_form_element_tangents(xb_model*);
int xb_assemble_tangent(xb_model*, double* A);
int xb_form_element_resids(xb_model*);
int xb_assemble_unbalance(xb_model*, double* B);
/* Interface exchange between the ranks of a partitioned model (NCCL send/recv over NVLink on
 * the model's stream); which = 0 element-tangent rows, 1 element-residual entries.
 * xb_form_tangent / xb_form_unbalance call it between their two halves.
 * xb_comm_unique_id: 128 bytes from ncclGetUniqueId on one rank, to be broadcast by the caller
 * (MPI / torch.distributed / a file); xb_comm_init: ncclCommInitRank(nparts, id, rank). */
int xb_comm_unique_id(char* out128);
int xb_comm_init(xb_model*, const char* id128);
int xb_exchange(xb_model*, int which);
/* the same exchange between n models (ranks 0..n-1) living in one process: plain device copies */
int xb_exchange_local(xb_model** models, int n, int which);
/* Domain::commit (Domain.cpp:1895) / Domain::revertToLastCommit (Domain.cpp:1925): the revert also puts the load
 * factor of the last commit back (currentTime = committedTime; applyLoad) and ends with an update */
int xb_commit(xb_model*);
int xb_revert_to_last_commit(xb_model*);
/* Domain::revertToStart (Domain.cpp:1951, the `reset` command): every node, material, section and element back to its
 * initial state, time and load factor 0, then an update; Element::Kc is left as it is (as in the reference) */
int xb_revert_to_start(xb_model*);
/* blocks until the model's stream is idle; surfaces asynchronous errors */
int xb_synchronize(xb_model*);

/* resident buffers (device pointers) for callers that keep the SOE on the GPU */
double* xb_device_A(xb_model*);
double* xb_device_B(xb_model*);
double* xb_device_trial_disp(xb_model*);

/* parity / debugging taps: Element::getTangentStiff / getResistingForce of FE element e
 * as last formed on the device (row-major nd*nd / nd doubles), and the material response
 * (stress [order], tangent [order*order]) of Gauss point g of element e */
int xb_get_element_tangent(xb_model*, long long e, double* K);
int xb_get_element_resid(xb_model*, long long e, double* R);
int xb_get_gp_response(xb_model*, long long e, int g, double* stress, double* tangent);

/* Run-time tuning of the device path; the results are bit-for-bit the same under every tuning setting.
 *   "ranged_tangent"  0 | 1 (default 1): run xb_form_tangent of a large single-batch brick model range by range on
 *                     two streams also when A stays on the device (with a host destination it always does)
 *   "brick_storage"   0 | 1 (default 1; before xb_setup): stdBrick tangents leave the tangent kernel as symmetric element
 *                     records (1: 324 doubles per element, gathered by the assembly) or as node-major rows (0: 576 doubles
 *                     per element, streamed by the assembly).  Measured on B200 at 4.1 M elements: records 13.2 ms per step
 *                     and 43.7 GB of DRAM traffic, rows 13.4 - 13.8 ms and 59 GB
 *   "tangent_ranges"  1..64 (default 8; before xb_setup): the number of element ranges of the ranged formTangent
 *   "fast_assembly"   0 | 1 (default 1): plain brick models (no MP constraints, rows <= 96 entries, <= 32 elements per
 *                     node) take the hand-tuned record assembly kernel; 0 forces the generic one
 * And one that follows the reference's `constraints` command (it changes results exactly as the command does there):
 *   "constraints_transformation"  0 | 1 (default 0 = `constraints Plain`; before xb_device_init): `constraints Transformation`
 *                     with fix / equalDOF constraints numbers and assembles like PlainHandler, but its enforceSPs() updates
 *                     every element with a constrained node once more at each applyLoad (analysis/handler/
 *                     TransformationConstraintHandler.cpp:462-483) -- after a commit that leaves a yielded J2 point with its
 *                     elastic tangent for the next step's first iteration, and makes a force-based beam iterate
 *                     once more from where it stood.  With 1, xb_apply_load repeats that update on the same elements
 * Returns XB_ERR_ARG for an unknown name or value. */
int xb_set_option(xb_model*, const char* name, int value);

/* kernel launches issued by this model since creation (bench.py's gpu_launches) */
long long xb_launch_count(const xb_model*);
/* bytes the last xb_form_tangent / xb_form_unbalance / xb_update moved algorithmically
 * (DESIGN.md "algorithmic bytes"): which = 0 update, 1 formUnbalance, 2 formTangent (compulsory
 * traffic of the whole operation), 3 element-tangent kernel, 4 tangent-assembly kernel,
 * 5 element-residual kernel (what each kernel must move given its inputs and outputs) */
long long xb_algorithmic_bytes(const xb_model*, int which);

#ifdef __cplusplus
}
#endif
#endif /* XARA_B200_H */
