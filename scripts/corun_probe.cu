// Development microbenchmark (B200): does other work on the same SM sub-partition overlap with FP64 DMMA?
// 8 warps per CTA, one CTA per SM: warps 0-3 (one per sub-partition) run DMMA chains with 8 accumulators; warps 4-7
// run `kind` = 0 nothing, 1 DFMA, 2 IMAD (int), 3 LDS.64, 4 FFMA.  Each role is timed with clock64 on its own warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256, 1) k(double* out, long long* clk, int iters_mma, int iters_other, int kind) {
  __shared__ double sm[4096];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  double s = 0;
  if (warp < 4) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters_mma; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) dmma884(c[j][0], c[j][1], a, b);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
  } else if (kind == 1) {
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters_other; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = fma(x[j], 1.0000001, 1e-9);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
  } else if (kind == 2) {
    int x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters_other; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = x[j] * 1664525 + 1013904223;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
  } else if (kind == 3) {
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters_other; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) s += sm[(idx + j * 32 + i) & 4095];
    }
  } else if (kind == 4) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters_other; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = fmaf(x[j], 1.0000001f, 1e-9f);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + warp] = t1 - t0;
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* d; cudaMalloc(&d, 1 << 24);
  long long* clk; cudaMalloc(&clk, 148 * 8 * sizeof(long long));
  const char* names[] = {"none", "DFMA", "IMAD", "LDS.64", "FFMA"};
  const int im = 4000, io = 4000;
  for (int kind = 0; kind < 5; kind++)
    for (int mma_on = 0; mma_on < 2; mma_on++) {
      if (kind == 0 && !mma_on) continue;
      k<<<148, 256>>>(d, clk, mma_on ? im : 0, io, kind);
      k<<<148, 256>>>(d, clk, mma_on ? im : 0, io, kind);
      cudaDeviceSynchronize();
      long long h[148 * 8]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double tm = 0, to = 0;
      for (int b = 0; b < 148; b++) for (int w = 0; w < 8; w++) (w < 4 ? tm : to) += h[b * 8 + w];
      tm /= 148 * 4; to /= 148 * 4;
      printf("other=%-6s dmma=%d : DMMA warp %.1f clk per DMMA (alone 16) | other warp %.2f clk per instruction  %s\n", names[kind], mma_on,
             mma_on ? tm / (im * 8.0) : 0.0, kind ? to / (io * 8.0) : 0.0, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
