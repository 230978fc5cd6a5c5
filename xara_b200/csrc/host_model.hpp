// Host-side mirror of the reference's Domain / AnalysisModel / LinearSOE set-up for
// the device path.  Integer work only: DOF_Group ids, FE_Element order, the sparse
// pattern and the per-node gather maps the assembly kernels consume.
//
// Reference counterparts (paths under /root/reference/SRC):
//   Domain (domain/domain/Domain.cpp)            -> HostModel node/element/SP/load tables
//   PlainHandler::handle (analysis/handler/PlainHandler.cpp:60)
//   PlainNumberer::numberDOF (analysis/numberer/PlainNumberer.cpp:73)
//   DOF_Numberer::numberDOF + RCM::number (analysis/numberer/DOF_Numberer.cpp:92,
//                                          graph/numberer/RCM.cpp:66)
//   AnalysisModel::getDOFGraph (analysis/model/AnalysisModel.cpp:286)
//   SparseGenColLinSOE::setSize / SparseGenRowLinSOE::setSize
#pragma once
#include <cstdint>
#include <string>
#include <vector>

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

// one xb_add_elements call: a homogeneous batch (one element kind, one material kind)
struct Group {
  int kind = 0, mat_kind = 0;
  std::vector<int> tag;      // [n]
  std::vector<int> conn;     // [n][nen] node TAGS until setup(), node indices after
  std::vector<int> mat;      // [n] material index into HostModel::mats
  std::vector<double> par;   // [n][npar]
  long long ke_off = 0;      // offset (doubles) of this group's element matrices in the Ke buffer
  long long re_off = 0;      // offset (doubles) of this group's element residuals
  long long gp_off = 0;      // first Gauss point (global numbering, for reporting)
  long long n() const { return (long long)tag.size(); }
};

struct HostModel {
  int ndm = 0, ndf = 0;
  // --- Domain ---
  std::vector<int> node_tag;        // ascending after setup()
  std::vector<double> crd;          // [nn][ndm]
  std::vector<int> sp_node, sp_dof; // fix
  std::vector<Material> mats;
  std::vector<Group> groups;
  std::vector<int> load_node;       // pending nodal loads (tags)
  std::vector<double> load_val;     // [nload][ndf]
  std::vector<double> load;         // [nn][ndf] after setup()

  // --- analysis (valid after setup) ---
  bool is_setup = false;
  int numberer = 0, soe_kind = 0;
  int neq = 0;
  std::vector<int> id;              // [nn][ndf] DOF_Group ids
  long long ne = 0;                 // FE_Elements
  std::vector<int> fe_group;        // [ne] group of FE element e (FE order = ascending element tag)
  std::vector<int> fe_local;        // [ne] index within the group
  std::vector<long long> ptr;       // [neq+1] colStartA / rowStartA
  std::vector<int> idx;             // [nnz]   rowA / colA
  // node -> FE elements, FE order within a node (this IS the reference's addA/addB
  // accumulation order for every entry owned by the node's equations)
  std::vector<long long> n2e_ptr;   // [nn+1]
  std::vector<long long> n2e_koff;  // [*] offset in Ke of row (a*ndf) of that element's matrix
  std::vector<long long> n2e_roff;  // [*] offset in Re of entry (a*ndf)
  std::vector<uint8_t> n2e_nd;      // [*] nd = nen*ndf of that element
  std::vector<int> n2e_fe;          // [*] FE index
  std::vector<uint8_t> n2e_loc;     // [*] local node a
  // node-level column lists: equations coupled to node n (sorted) = the pattern of each of
  // its free dofs' rows/columns
  std::vector<long long> ncol_ptr;  // [nn+1]
  std::vector<int> ncol;            // [*]
  int cp_stride = 0;                // max nd over groups
  std::vector<uint16_t> colpos;     // [n2e_total][cp_stride]: position of local dof j's equation
                                    //   inside node n's column list, 0xFFFF if constrained
  int max_row = 0;                  // longest row
  long long ke_total = 0, re_total = 0, ngp = 0;

  std::string err;

  int add_nodes(int n, const int* tags, const double* c);
  int add_sp(int n, const int* tags, const int* dofs);
  int add_material(int tag, int kind, const double* par, int npar);
  int add_elements(int kind, int n, const int* tags, const int* conn, const int* mat_tags,
                   const double* par, int par_stride);
  int add_loads(int n, const int* tags, const double* vals);
  int setup(int numberer, int soe_kind);
  long long nnz() const { return ptr.empty() ? 0 : ptr.back(); }
  int nn() const { return (int)node_tag.size(); }
  // addA location of element-matrix entries, FE elements [e0,e1)
  int scatter_map(long long e0, long long e1, long long* map) const;
};

}  // namespace xb
