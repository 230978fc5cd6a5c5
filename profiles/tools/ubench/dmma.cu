// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) rate on this GPU against the plain DFMA rate, and the brick tangent's
// per-Gauss-point contraction K(24x24) += B^T(24x6) . (D B)(6x24) written both ways:
//   dmma: upper-triangular 8x8 tiles (6 of 9), k = 6 padded to 8 -> 12 mma.m8n8k4 per Gauss point, 8 points
//   dfma: the rank-1 form of brick_tangent (219 FP64 instructions per lane and Gauss point, 8 lanes per element)
// Operands are synthetic registers: this is the arithmetic ceiling of either form, nothing else.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int C>   // C independent accumulator tiles per warp
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c[C][2];
#pragma unroll
  for (int i = 0; i < C; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < C; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; i++) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}

template <int C>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[C];
#pragma unroll
  for (int i = 0; i < C; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < C; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; i++) s += x[i];
  if (s == 1.2345) out[0] = s;
}

// one element per warp and iteration: 8 Gauss points x 6 upper tiles x 2 k-steps, operands rotated so nothing folds
__global__ void k_contract_dmma(double* out, int nelem_per_warp, double a0, double b0) {
  double acc = 0;
  for (int e = 0; e < nelem_per_warp; e++) {
    double c[6][2];
#pragma unroll
    for (int t = 0; t < 6; t++) { c[t][0] = 0; c[t][1] = 0; }
    double a = a0 + e, b = b0 + threadIdx.x;
#pragma unroll 1
    for (int g = 0; g < 8; g++) {
      // row tiles R0..R2 (a operands), column tiles C0..C2 (b operands), 2 k-steps each: 12 operand registers per point
      double ar[3][2], bc[3][2];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int kk = 0; kk < 2; kk++) { ar[i][kk] = a + i + 0.5 * kk + g; bc[i][kk] = b - i + 0.25 * kk + g; }
      int t = 0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++, t++) { dmma(c[t][0], c[t][1], ar[i][0], bc[j][0]); dmma(c[t][0], c[t][1], ar[i][1], bc[j][1]); }
    }
#pragma unroll
    for (int t = 0; t < 6; t++) acc += c[t][0] + c[t][1];
  }
  if (acc == 1.2345) out[0] = acc;
}

// the rank-1 DFMA form: 8 lanes per element, 4 elements per warp and iteration; per lane and point 5 blocks x 33 + 54
__global__ void k_contract_dfma(double* out, int nbatch_per_warp, double a0, double b0) {
  double tot = 0;
  for (int e = 0; e < nbatch_per_warp; e++) {
    double acc[5][3][3];
#pragma unroll
    for (int t = 0; t < 5; t++)
#pragma unroll
      for (int p = 0; p < 3; p++)
#pragma unroll
        for (int q = 0; q < 3; q++) acc[t][p][q] = 0;
#pragma unroll 1
    for (int g = 0; g < 8; g++) {
      const double gk[3] = {a0 + g, a0 - g, a0 + 2 * g + threadIdx.x}, ca = b0 + g, cb = b0 - g, cg = b0 + e;
      const double n[6] = {ca, cb, cg, ca + 1, cb + 1, cg + 1};
      double ak[3], bk[3], vk[3], wk[3];
#pragma unroll
      for (int p = 0; p < 3; p++) { ak[p] = ca * gk[p]; bk[p] = cb * gk[p]; }
      vk[0] = gk[0] * n[0] + gk[1] * n[3] + gk[2] * n[5]; vk[1] = gk[1] * n[1] + gk[0] * n[3] + gk[2] * n[4];
      vk[2] = gk[2] * n[2] + gk[1] * n[4] + gk[0] * n[5];
#pragma unroll
      for (int p = 0; p < 3; p++) wk[p] = cg * vk[p];
#pragma unroll
      for (int t = 0; t < 5; t++) {
        const double gJ[3] = {gk[0] + t, gk[1] - t, gk[2] + 0.5 * t};
        double vJ[3];
        vJ[0] = gJ[0] * n[0] + gJ[1] * n[3] + gJ[2] * n[5]; vJ[1] = gJ[1] * n[1] + gJ[0] * n[3] + gJ[2] * n[4];
        vJ[2] = gJ[2] * n[2] + gJ[1] * n[4] + gJ[0] * n[5];
        const double sd = fma(gJ[2], bk[2], fma(gJ[1], bk[1], gJ[0] * bk[0]));
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
          for (int q = 0; q < 3; q++) acc[t][p][q] = fma(vJ[p], wk[q], fma(bk[p], gJ[q], fma(gJ[p], ak[q], acc[t][p][q])));
        acc[t][0][0] += sd; acc[t][1][1] += sd; acc[t][2][2] += sd;
      }
    }
#pragma unroll
    for (int t = 0; t < 5; t++)
#pragma unroll
      for (int p = 0; p < 3; p++)
#pragma unroll
        for (int q = 0; q < 3; q++) tot += acc[t][p][q];
  }
  if (tot == 1.2345) out[0] = tot;
}

template <class F>
static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* d; cudaMalloc(&d, 8);
  printf("%s, %d SMs\n", p.name, sms);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {
    const int blocks = sms * wps / 4;
    float ms = timeit([&] { k_dfma<8><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9); });
    printf("DFMA  8 chains, %2d warps/SM: %6.2f T FMA/s\n", wps, (double)blocks * 128 * 8 * iters / ms / 1e9);
    ms = timeit([&] { k_dmma<2><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9); });
    printf("DMMA  2 tiles,  %2d warps/SM: %6.2f T FMA/s (%.3f G mma/s)\n", wps, (double)blocks * 4 * 2 * iters * 256 / ms / 1e9, (double)blocks * 4 * 2 * iters / ms / 1e6);
    ms = timeit([&] { k_dmma<6><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9); });
    printf("DMMA  6 tiles,  %2d warps/SM: %6.2f T FMA/s (%.3f G mma/s)\n", wps, (double)blocks * 4 * 6 * iters * 256 / ms / 1e9, (double)blocks * 4 * 6 * iters / ms / 1e6);
  }
  for (int wps : {8, 16, 32}) {
    const int blocks = sms * wps / 4, per = 2000;
    float ms = timeit([&] { k_contract_dmma<<<blocks, 128>>>(d, per, 1.0000001, 1e-9); });
    const double ne1 = (double)blocks * 4 * per;
    printf("contraction DMMA (12 mma/point, symmetric tiles), %2d warps/SM: %7.1f M elements/s  -> 4.096 M elements in %.2f ms\n", wps, ne1 / ms / 1e3, 4.096e6 / (ne1 / ms));
    ms = timeit([&] { k_contract_dfma<<<blocks, 128>>>(d, per, 1.0000001, 1e-9); });
    const double ne2 = (double)blocks * 4 * per * 4;
    printf("contraction DFMA (rank-1 form, 8 lanes/element),  %2d warps/SM: %7.1f M elements/s  -> 4.096 M elements in %.2f ms\n", wps, ne2 / ms / 1e3, 4.096e6 / (ne2 / ms));
  }
  return 0;
}
