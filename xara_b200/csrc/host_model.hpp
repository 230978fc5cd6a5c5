// Host-side mirror of the reference's Domain / AnalysisModel / LinearSOE set-up for
// the device path.  Integer work only: DOF_Group ids, FE_Element order, the sparse
// pattern and the per-node gather maps the assembly kernels consume; for multi-GPU
// runs also the element partition, node ownership and the interface exchange lists.
//
// Reference counterparts (paths under /root/reference/SRC):
//   Domain (domain/domain/Domain.cpp)            -> HostModel node/element/SP/load tables
//   PlainHandler::handle (analysis/handler/PlainHandler.cpp:60)
//   PlainNumberer::numberDOF (analysis/numberer/PlainNumberer.cpp:73)
//   DOF_Numberer::numberDOF + RCM::number (analysis/numberer/DOF_Numberer.cpp:92,
//                                          graph/numberer/RCM.cpp:66)
//   AnalysisModel::getDOFGraph (analysis/model/AnalysisModel.cpp:286)
//   SparseGenColLinSOE::setSize / SparseGenRowLinSOE::setSize
//   DomainPartitioner / PartitionedDomain (domain/partitioner, domain/domain/partitioned)
//       -> element partition + row ownership; the numbering stays the GLOBAL one
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "brick_rec.hpp"

namespace xb {

struct EleKind {
  int nen;   // nodes per element
  int ndf;   // dofs per node the element uses
  int nip;   // integration points
  int nst;   // stress components per point
  int npar;  // element parameters kept on the device
};
const EleKind& ele_kind(int kind);

struct Material {
  int tag, kind;
  double par[8];
};

struct Uniaxial {
  int tag, kind;
  double par[12];
};
// section Fiber (FiberSection2d): fibres in the order they were added
struct FiberSectionDef {
  int tag = 0;
  std::vector<double> y, A;
  std::vector<int> mat;      // index into HostModel::unis
  double yBar = 0.0;         // FiberSection2d::addFiber keeps QzBar/ABar up to date
  // section Aggregator of two uniaxial materials (P, Mz): "fibre" 0 takes the axial strain, "fibre" 1 the curvature
  bool agg = false;
  // section Fiber -GJ in a 3D model (FiberSection3d): fibre z, centroid zBar, elastic torsion GJ
  bool is3d = false;
  std::vector<double> z;
  double zBar = 0.0, GJ = 0.0;
};

// one xb_add_elements call: a homogeneous batch (one element kind, one material kind).
// After setup() it holds only the elements of this rank's partition.
struct Group {
  int kind = 0, mat_kind = 0;
  std::vector<int> tag;      // [n]
  std::vector<int> conn;     // [n][nen] node TAGS until setup(), LOCAL node indices after
  std::vector<int> mat;      // [n] material index into HostModel::mats
  std::vector<double> par;   // [n][npar]
  // forceBeamColumn batches: one fibre section, nIP Lobatto points, element iteration controls
  int sec = -1, nip = 0, max_iters = 10;
  int transf = 0;            // geomTransf of the batch: 0 Linear, 1 PDelta, 2 Corotational (2D)
  std::vector<double> rule;  // forceBeamColumn: per element nip locations then nip weights (xb_set_beam_integration); empty: Lobatto
  std::vector<uint8_t> rule_set;
  double tol = 1e-12;
  bool j2_plane_stress = false;   // FourNodeQuad batch whose J2Plasticity copies are J2PlaneStress
  std::vector<long long> kdst;  // [n][nen] where the rows of node a of element l go: >= 0 offset of the
                                //   slot in the node-major buffer KeN (node owned here), < 0 offset
                                //   -(x+1) in the send buffer (node owned by another rank).  Empty for stdBrick
                                //   batches: their tangents are kept as symmetric element records (rec_off)
  long long rec_off = 0;     // stdBrick: offset (doubles) of the batch's first element record (brick_rec.hpp)
  long long ke_off = 0;      // (legacy element-major offset; sizes the residual bookkeeping)
  long long re_off = 0;      // offset (doubles) of this group's element residuals
  long long gp_off = 0;      // first Gauss point (for reporting)
  long long n() const { return (long long)tag.size(); }
};

// one neighbour rank in the interface exchange
struct Peer {
  int rank = 0;
  long long send_k = 0, recv_k = 0;   // doubles of element-matrix rows sent to / received from it
  long long send_r = 0, recv_r = 0;   // doubles of element-residual entries
  long long send_k_base = 0, recv_k_base = 0, send_r_base = 0, recv_r_base = 0;  // offsets in the buffers
  long long chunks_out = 0, chunks_in = 0;
};

struct HostModel {
  int ndm = 0, ndf = 0;
  // --- Domain (global until setup(), then restricted to this rank's local nodes) ---
  std::vector<int> node_tag;        // ascending
  std::vector<double> crd;          // [nn][ndm]
  std::vector<int> sp_node, sp_dof; // fix
  // nodes created under another `model -ndf` (a FourNodeQuad's 2-dof nodes beside 3-dof frame nodes, FourNodeQuad.cpp:
  // 133-139): (tag, dofs) pairs; every other node has the model's ndf.  The dofs a node does not have are carried as
  // constrained (-1): they get no equation, exactly as the node's shorter DOF_Group::myID in the reference
  std::vector<int> ndf_node, ndf_val;
  int set_node_ndf(int n, const int* tags, int nd);
  std::vector<Material> mats;
  std::vector<Uniaxial> unis;
  std::vector<FiberSectionDef> secs;
  std::vector<Group> groups;
  std::vector<int> load_node;       // pending nodal loads (tags)
  std::vector<double> load_val;     // [nload][ndf]
  std::vector<double> load;         // [nn][ndf]
  std::vector<int> mass_node;       // pending `mass` commands (tags)
  std::vector<double> mass_val;     // [n][ndf]
  std::vector<double> mass;         // [nn][ndf] diagonal of Node::mass
  std::vector<uint16_t> diagpos;    // [nn][ndf] position of the dof's own equation in the node's column list

  // --- analysis (valid after setup) ---
  bool is_setup = false;
  int numberer = 0, soe_kind = 0;   // soe_kind: the pattern's orientation (SparseGenCol | SparseGenRow)
  // the SOE whose A is filled (XB_SOE_*): the compressed ones, or BandGeneral / ProfileSPD, which keep the compressed-
  // column pattern for the assembly and map every pattern entry to its location in their own array (a_loc)
  int soe_store = 0;
  std::vector<long long> a_loc;     // [nnz] location in A of pattern entry k, -1: not stored (ProfileSPD lower triangle); empty: k itself
  long long a_total = 0;            // length of A
  int band_sub = 0, band_super = 0; // BandGenLinSOE::numSubD / numSuperD
  std::vector<int> profile_diag;    // ProfileSPDLinSOE::iDiagLoc (1-based)
  long long a_size() const { return a_loc.empty() ? nnz() : a_total; }
  int nparts = 1, rank = 0;
  int neq = 0;                      // GLOBAL number of equations
  int nn_global = 0;
  long long ne_global = 0;
  std::vector<int> id;              // [nn][ndf] GLOBAL DOF_Group ids of the local nodes
  std::vector<int> row_of;          // [nn][ndf] local row of an owned free dof, else -1
  // ---- MP constraints (`equalDOF`): PlainHandler's -4 ids.  A tied dof shares the retained dof's equation, so
  // that row gathers from the elements of several nodes and has its own column list: "shared rows" are kept out
  // of the per-node assembly tasks (row_of_dev = -1, task row offset = -1) and assembled by row (irr_*).
  std::vector<int> mp_r, mp_c, mp_dof;       // retained / constrained node tag, dof
  std::vector<int> row_of_dev;               // row_of with the dofs of shared equations set to -1
  int max_dup = 0;                           // most element dofs of ONE element on one equation, minus 1 (colpos rank bits)
  int irr_max_row = 0;
  std::vector<int> irr_row;                  // [nirr] local row
  std::vector<long long> irr_ptr;            // [nirr+1] -> entries, in (FE_Element, element dof) order = addA / addB order
  std::vector<long long> irr_src;            // [*] offset of the element-matrix row in KeN
  std::vector<long long> irr_roff;           // [*] offset of the element-residual entry in Re
  std::vector<uint16_t> irr_cp;              // [*][cp_stride] column positions in the row's own list (+ rank bits)
  std::vector<long long> irr_own_ptr;        // [nirr+1] -> owners
  std::vector<int> irr_own;                  // [*] node * ndf + dof of every (node, dof) on the equation, DOF_Group order
  std::vector<uint16_t> irr_diag;            // [nirr] position of the equation in its own row
  int add_equal_dof(int r_tag, int c_tag, int n, const int* dofs);
  std::vector<uint8_t> owned;       // [nn] this rank owns the node's equations
  int nrows = 0;                    // owned equations (== neq when nparts == 1)
  std::vector<int> row_geq;         // [nrows] their global numbers, ascending
  long long ne = 0;                 // local FE_Elements
  std::vector<int> fe_group;        // [ne] group of local FE element e (FE order = ascending element tag)
  std::vector<int> fe_local;        // [ne] index within the group
  std::vector<long long> fe_global; // [ne] position in the global FE_Element order
  std::vector<long long> ptr;       // [nrows+1] colStartA / rowStartA of the owned rows
  std::vector<int> idx;             // [nnz]     rowA / colA (GLOBAL equation numbers)
  // node -> FE elements, GLOBAL FE order within a node (this IS the reference's addA/addB
  // accumulation order for every entry owned by the node's equations).  Only owned nodes
  // carry slots.  Element-tangent rows are stored NODE-MAJOR: slot t of the list occupies
  // KeN[t*chunk ...] (ndf rows x cp_stride columns), written directly by the element kernel
  // (or copied from the receive buffer when the element lives on another rank), so a node's
  // rows are one contiguous stream.  Residual entries of remote elements are read from the
  // receive buffer: roff < 0 encodes offset -(x+1).
  std::vector<long long> n2e_ptr;   // [nn+1]
  std::vector<long long> n2e_koff;  // [*] offset of the slot in KeN (= slot index * chunk)
  std::vector<long long> n2e_roff;  // [*] offset in Re of entry (a*ndf)
  // record models (stdBrick, see brick_rec.hpp): where the slot's rows come from.  offset << 4 | local node << 1 | 1:
  // gathered from the element record at `offset` (doubles); offset << 4 | 0: ndf dense rows of cp_stride columns at
  // `offset` of the receive buffer (the element lives on another rank).  Empty otherwise (slot u = KeN[u * chunk]).
  std::vector<long long> n2e_ksrc;
  bool rec_mode = false;            // every batch is stdBrick and brick_records is on: element records + gathered assembly
  bool brick_records = true;        // stdBrick tangents as symmetric element records (25 % less HBM traffic and 8 GB less at 4 M
                                    //   elements, assembly bound by the L1 data pipe) instead of node-major rows (streamed by the
                                    //   assembly at the HBM rate) -- xb_set_option "brick_storage" before xb_setup
  long long rec_total = 0;          // doubles of element records
  std::vector<long long> pk_src;    // record models: outgoing row chunk c (send buffer offset c * chunk) <- record descriptor
  // the same descriptors in 32 bits (offset in units of 36 doubles << 4 | local node, 8 = dense rows) for the hand-tuned
  // assembly kernel, which takes plain brick models: no MP constraints, rows of at most 96 entries, <= 32 elements per node
  std::vector<uint32_t> n2e_ksrc32;
  bool fast_asm_ok = false;
  std::vector<uint8_t> n2e_nd;      // [*] nd = nen*ndf of that element
  std::vector<long long> n2e_fe;    // [*] GLOBAL FE index
  std::vector<uint8_t> n2e_loc;     // [*] local node a
  // node-level column lists: equations coupled to node n (sorted) = the pattern of each of
  // its free dofs' rows/columns
  std::vector<long long> ncol_ptr;  // [nn+1]
  std::vector<int> ncol;            // [*]
  int cp_stride = 0;                // max nd over groups
  std::vector<uint16_t> colpos;     // [n2e_total][cp_stride]: position of local dof j's equation
                                    //   inside node n's column list, 0xFFFF if constrained
  // ranged formTangent: the elements are cut into `nchunk` consecutive ranges (per batch); a
  // node can be assembled once the last range holding one of its elements is done.  node_perm
  // lists the owned nodes by that range (nodes with rows from other ranks last: range nchunk);
  // record models: inside a range in Morton order of the node coordinates, so that the records a
  // node gathers from are still in L2 when its neighbours ask for them.
  int want_ranges = 8;              // element ranges of the ranged formTangent (xb_set_option "tangent_ranges" before xb_setup)
  int nchunk = 1;
  std::vector<int> node_perm;       // [n_owned_nodes_with_rows]
  std::vector<long long> chunk_node_ptr;   // [nchunk+2]
  std::vector<long long> asm_task;         // [node_perm.size()][3+ndf] per-node record of the assembly kernel
  std::vector<long long> chunk_a_ptr;      // [nchunk+2] offsets in A of the rows each range completes
  bool rows_streamable = false;     // the ranges own consecutive row blocks of A (copy-out can follow them)
  int max_row = 0;                  // longest row
  long long ke_total = 0, re_total = 0, ngp = 0;
  int chunk = 0;                    // doubles per slot: ndf rows x cp_stride columns
  long long kn_total = 0;           // doubles in KeN (node-major element rows of the owned nodes)
  std::vector<long long> uk_src, uk_dst;   // received matrix chunk -> its slot in KeN

  // --- interface exchange (nparts > 1) ---
  std::vector<Peer> peers;          // ascending rank
  std::vector<long long> pr_src, pr_dst;   // per outgoing residual chunk: offset in Re, offset in the send buffer
  long long send_k_total = 0, recv_k_total = 0, send_r_total = 0, recv_r_total = 0;
  std::vector<int> part_fe;         // [ne_global] partition of every element (global FE order)

  std::string err;

  int add_nodes(int n, const int* tags, const double* c);
  int add_sp(int n, const int* tags, const int* dofs);
  int add_material(int tag, int kind, const double* par, int npar);
  int add_uniaxial(int tag, int kind, const double* par, int npar);
  int add_fiber_section(int tag, int nf, const double* y, const double* A, const int* mat_tags);
  int add_section_aggregator(int tag, int n, const int* mat_tags, const int* codes);
  int add_fiber_section3d(int tag, int nf, const double* y, const double* z, const double* A, const int* mat_tags, double GJ);
  int add_elements(int kind, int n, const int* tags, const int* conn, const int* mat_tags,
                   const double* par, int par_stride);
  int add_loads(int n, const int* tags, const double* vals);
  int add_beam_uniform_loads(int n, const int* tags, const double* w);
  int add_beam_point_loads(int n, const int* tags, const double* p);
  int add_beam_partial_loads(int n, const int* tags, const double* p);
  int set_beam_integration(int n, const int* tags, int nip, const double* xi, const double* wt);
  int add_mass(int n, const int* tags, const double* vals);
  // part: nullptr (built-in recursive coordinate bisection) or [ne] ranks in FE order
  int setup(int numberer, int soe_kind, int nparts = 1, int rank = 0, const int* part = nullptr);
  long long nnz() const { return ptr.empty() ? 0 : ptr.back(); }
  int nn() const { return (int)node_tag.size(); }
  // addA location of element-matrix entries, local FE elements [e0,e1); -1 = not owned here / dropped
  int scatter_map(long long e0, long long e1, long long* map) const;
};

}  // namespace xb
