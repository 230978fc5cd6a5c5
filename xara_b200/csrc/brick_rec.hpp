// Symmetric element record of a stdBrick tangent (host + device).
//
// The 24x24 element tangent of the J2 / elastic brick is symmetric node-pair block by node-pair block
// (K_kJ = K_Jk^T; both orientations are the same FP64 values), so a brick's tangent is kept as the 36 distinct 3x3
// blocks = 324 doubles, element-major, instead of 8 node slots of 3 x 24 (576 doubles).  The tangent kernel's lane k
// (k = local node) forms the blocks K_Jk, J = k, k+1, .., k+4 (mod 8; lanes 4-7 stop at k+3): region k of the record
// holds them one after the other,
//     record[region(k) + 9 t + 3 a + b] = K[(J = (k+t) & 7, dof a)][(k, dof b)],   t = 0..4 (k < 4), 0..3 (k >= 4)
// with the 45- and 36-double regions interleaved (0 4 1 5 2 6 3 7) so that the lanes' shared-memory writes spread
// over the banks.  The assembly gathers the rows of local node J out of the record: K[(J,p)][(K,q)] sits in region K
// when J - K (mod 8) is one of K's offsets, else transposed in region J.
#pragma once
#ifdef __CUDACC__
#define XB_REC_HD __host__ __device__ __forceinline__
#else
#define XB_REC_HD inline
#endif

namespace xb {

constexpr int kBrickRec = 324;

XB_REC_HD constexpr int brick_rec_region(int k) { return (k & 3) * 81 + (k >> 2) * 45; }

// offset in the record of entry ((J,p),(K,q)) of the STORED matrix: the element tangent for a row-compressed SOE,
// its transpose for a column-compressed one (transpose != 0), as FE_Element hands it to addA
XB_REC_HD constexpr int brick_rec_entry(int J, int p, int K, int q, int transpose) {
  if (transpose) { const int tJ = J, tp = p; J = K; p = q; K = tJ; q = tp; }
  const int dt = (J - K) & 7;                       // J = K + dt: block (J, K) of lane K, if it forms that offset
  if (dt < 4 || (dt == 4 && K < 4)) return brick_rec_region(K) + 9 * dt + 3 * p + q;
  return brick_rec_region(J) + 9 * ((K - J) & 7) + 3 * q + p;   // block (K, J) of lane J, transposed
}

// slot descriptor of the gathered assembly (HostModel::n2e_ksrc)
XB_REC_HD constexpr long long brick_rec_desc(long long offset, int local_node) { return (offset << 4) | ((long long)local_node << 1) | 1; }
XB_REC_HD constexpr long long dense_rows_desc(long long offset) { return offset << 4; }

}  // namespace xb
