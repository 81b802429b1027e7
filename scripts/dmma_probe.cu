// Microbenchmark: FP64 DMMA (mma.sync m8n8k4) throughput vs resident warps per SM and independent accumulators per warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int j = 0; j < NACC; j++) c[j][0] = c[j][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) dmma884(c[j][0], c[j][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j][0] + c[j][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps_per_sm, int nsm, double* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<NACC><<<nsm, warps_per_sm * 32>>>(d, 100);
  cudaEventRecord(e0);
  k<NACC><<<nsm, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flop = (double)nsm * warps_per_sm * iters * NACC * 512.0;
  double clk_per_dmma_per_sm = (ms * 1e-3 * 1.965e9) / ((double)warps_per_sm * iters * NACC);
  printf("warps/SM %2d  acc/warp %d : %6.2f TFLOP/s   %.2f clk per DMMA per SM, %.1f clk per DMMA per warp\n", warps_per_sm, NACC,
         flop / (ms * 1e-3) / 1e12, clk_per_dmma_per_sm, clk_per_dmma_per_sm * warps_per_sm);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 1 << 24);
  for (int w : {1, 2, 4, 6, 8, 12, 16, 32}) { run<1>(w, p.multiProcessorCount, d); run<2>(w, p.multiProcessorCount, d); run<4>(w, p.multiProcessorCount, d); run<8>(w, p.multiProcessorCount, d); }
  return 0;
}
