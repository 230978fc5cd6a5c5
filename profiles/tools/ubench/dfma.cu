// FP64 FMA throughput on this GPU: W warps per SM, C independent chains per thread
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, double a, double b) {
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
template <int C>
void run(int warps_per_sm, int sms) {
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const int threads = 128, blocks = sms * warps_per_sm / 4;
  k<C><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<C><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * threads * C * iters;
  printf("chains %2d warps/SM %2d: %.2f T DFMA/s  (%.1f DFMA/clk/SM at 1.9 GHz)\n", C, warps_per_sm, fma / ms / 1e9, fma / ms / 1e6 / sms / 1.9e3);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<1>(8, sms); run<2>(8, sms); run<4>(8, sms); run<8>(8, sms); run<16>(8, sms);
  run<4>(4, sms); run<8>(4, sms); run<16>(4, sms);
  run<8>(16, sms); run<8>(32, sms); run<4>(64, sms);
  return 0;
}
