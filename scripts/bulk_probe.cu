// Development microbenchmark (B200): throughput of 1-D bulk copies (cp.async.bulk global -> shared, mbarrier
// completion) per SM as a function of copy size, copies per batch and batches in flight, from an L2-resident table.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe scripts/bulk_probe.cu ; run: ./bulk_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned ok;
  do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// each batch = ncopy copies of `size` bytes into one stage; `depth` stages in flight; nbatch batches per CTA
__global__ void __launch_bounds__(128, 1) probe(const char* src, size_t footprint, int size, int ncopy, int depth, int nbatch, long long* clk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar[32];
  const int stage_bytes = size * ncopy;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long seed = blockIdx.x * 7919ull + 12345ull;
    auto issue = [&](int b) {
      const int st = b % depth;
      mbar_expect_tx(&bar[st], (unsigned)stage_bytes);
      for (int c = 0; c < ncopy; c++) {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        const size_t off = ((seed >> 20) & (footprint / 256 - 1)) * 128;   // footprint is a power of two; stays in the lower half + size
        bulk_g2s(smem + (size_t)st * stage_bytes + (size_t)c * size, src + off, (unsigned)size, &bar[st]);
      }
    };
    const long long t0 = clock64();
    for (int b = 0; b < depth && b < nbatch; b++) issue(b);
    for (int b = 0; b < nbatch; b++) {
      mbar_wait(&bar[b % depth], (b / depth) & 1);
      if (b + depth < nbatch) issue(b + depth);
    }
    clk[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t footprint = 64ull << 20;   // L2-resident on B200 (126 MB L2)
  char* src; cudaMalloc(&src, footprint); cudaMemset(src, 1, footprint);
  long long* clk; cudaMalloc(&clk, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int configs[][3] = {  // size, ncopy, depth
      {1024, 16, 1}, {1024, 16, 2}, {1024, 16, 4}, {1024, 16, 8}, {4096, 4, 1}, {4096, 4, 2}, {4096, 4, 4}, {4096, 4, 8},
      {4096, 10, 1}, {4096, 10, 2}, {4096, 10, 4}, {16384, 1, 1}, {16384, 1, 2}, {16384, 1, 4}, {16384, 1, 8}, {16384, 4, 1},
      {16384, 4, 2}, {16384, 4, 3}, {32768, 2, 1}, {32768, 2, 2}, {32768, 2, 3}, {65536, 1, 1}, {65536, 1, 2}, {65536, 1, 3},
      {128, 64, 4}, {256, 64, 4}, {512, 32, 4}, {2048, 32, 2}, {8192, 8, 2}, {8192, 3, 4}, {24576, 3, 2}, {24576, 1, 6}};
  for (auto& c : configs) {
    const int size = c[0], ncopy = c[1], depth = c[2];
    const int nbatch = (int)((64ull << 20) / ((size_t)size * ncopy * 148) + 8);
    const size_t smem = (size_t)size * ncopy * depth;
    if (smem > 200 * 1024) continue;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<148, 128, smem>>>(src, footprint, size, ncopy, depth, nbatch, clk);   // warm L2
    cudaEventRecord(e0);
    probe<<<148, 128, smem>>>(src, footprint, size, ncopy, depth, nbatch, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double bytes = (double)size * ncopy * nbatch;
    printf("size %6d x %2d copies/batch, depth %d (%6.1f KB in flight): %6.2f TB/s aggregate, %5.1f B/clk/SM, %7.0f clk/batch  err=%s\n",
           size, ncopy, depth, smem / 1024.0, bytes * 148 / (ms * 1e-3) / 1e12, bytes / avg, avg / nbatch, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
