// Shared device-side definitions of the B200 FAM iteration (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define PNFAM_CUDA_CHECK(x)                                                                          \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess)                                                                           \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                            \
  } while (0)

namespace pnfam {

// ---- tiling constants -------------------------------------------------------------------------
constexpr int RT = 16;        // grid points per r-tile (2 DMMA m-tiles)
constexpr int RS = 20;        // padded r-stride of a wave-function row in shared memory (bank-conflict free)
constexpr int NTYPE = 5;      // wf, d/dr, (Lambda/r), d/dz, laplacian_all

// Padded index space: inside every block the spin-up states and the spin-down states are each padded to a multiple
// of 4 rows (zero wave functions), so that DMMA k-steps / n-tiles never straddle a spin boundary.  pstart[block] is the
// first padded row of a block; a chunk of consecutive padded rows is what the kernels stage in shared memory.
// Wave-function tables (all rows in the padded space, all chunk copies linear = ONE cp.async.bulk each):
//   phi4[r-tile][row][4 types][RT]   density (rho):   wf, d/dr, Lambda/r, d/dz
//   phi5[r-tile][row][5 types][RT]   projection (h):  the same + laplacian
//   phi0[4-tile group][row][4][RT]   kappa / Delta:   wf of four consecutive r-tiles
// Inside each 128-byte (row, slot) line the 16 grid points are rotated by phi_rot(row) positions: 0, 8, 4, 12 for
// row & 3 = 0, 1, 2, 3.  Lines copied linearly into shared memory are then bank-conflict free without padding for
// both access patterns of the kernels: DMMA fragments (4 consecutive rows x 4 consecutive points -> 16 different
// 8-byte banks) and the G build (8 consecutive points of two adjacent rows).  Chunks start at multiples of 4 rows, so
// the rotation of a lane's rows is a function of its lane index alone.
__host__ __device__ __forceinline__ int phi_rot(int state) { return ((state & 1) << 3) | ((state & 2) << 1); }

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2*(l%4)] and C[l/4][2*(l%4)+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier + bulk-copy primitives (PTX) --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  const unsigned a = smem_u32(b);
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool mbar_test(unsigned long long* b, unsigned parity) {   // non-blocking
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

struct cplx {
  double re, im;
};
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return {-a.re, -a.im}; }
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return {s * a.re, s * a.im}; }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__host__ __device__ __forceinline__ cplx mul_i(cplx a) { return {-a.im, a.re}; }     // i * a
__host__ __device__ __forceinline__ cplx mul_mi(cplx a) { return {a.im, -a.re}; }    // -i * a

// ---- block structures on the device ---------------------------------------------------------------
struct DevBlockStruct {
  const int* r2c;   // [nb] partner column block or -1
  const int* r2m;   // [nb] element offset of the block
};

struct DevBasis {
  int nb, dqp, nghl, ntiles;
  const int* db;        // [nb]
  const int* isstart;   // [nb] 0-based first state of block
  const int* nsu;       // [nb] number of spin-up states (they come first inside a block)
  const int* pstart;    // [nb] first row of the block in the padded index space
  int dqp_p;            // rows of the padded index space
  const double* phi4;   // wave-function tables, see above
  const double* phi5;
  const double* phi0;
  const double* wdcori; // [nghl]
  const double* crho;   // [nghl]
  const double* cs;
  const double* cpair;
  const double* cspair;
  double cdrho, ctau, ctj0, ctj1, ctj2, crdj, cds, ct, cj, cgs, cf, csdj;
};

// transform tasks (host/symbolic.hpp flattened)
struct DevTerm {
  int a_mat, a_off, a_trans;
  int b_quad, b_off, b_trans;
  int c_mat, c_off, c_trans;
  int t_off;                 // offset of the intermediate op(A)op(B) block in the scratch array
  double alpha_re, alpha_im;
};
struct DevTask {
  int out_quad, out_off, m, n, nterms;
  DevTerm t[4];
};


// fused transform jobs (transform.cu): the terms of one or two output blocks regrouped by associativity,
//   T_x = sum_{t in group x} alpha_t op(A_t) op(B_t)            (phase 1, the products two outputs share are formed once)
//   out_o = sum_x beta_{o,x} T_x op(C_{o,x})                    (phase 2, T_x stays in the registers of the warp)
struct FusedTerm {
  int a_mat, a_off, a_trans;
  int b_quad, b_off, b_trans;
  double alpha_re, alpha_im;
};
struct FusedOut {
  int out_quad, out_off;
  int c_mat[2], c_off[2], c_trans[2];     // per group
  double beta_re[2], beta_im[2];
};
struct FusedJob {
  int m, n, ngroups, nout;
  int nterms[2];
  FusedTerm t[2][2];                      // [group][term]
  FusedOut o[2];
};

}  // namespace pnfam
