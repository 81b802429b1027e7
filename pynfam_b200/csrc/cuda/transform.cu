// Blocks (a) and (d) of the FAM iteration: quasiparticle <-> single-particle transforms
//     dRsp = W_a . dRqp . W_b^T       (reference: pnfam_solver.f90:144-150 -> triprod_bbm -> 2 dgemm per block,
//     dHqp = W_a^T . dHsp . W_b        pnfam_solver.f90:160-166,       pnfam_type_blockmatrix.f90:202-206)
// executed as two grouped-GEMM launches over the task list built by host/symbolic.cpp:
//   phase 1:  T_t   = op(A_t) . op(B_t)               one CTA per (non-empty 32x32 tile of a task term, point, re/im)
//   phase 2:  out   = sum_t alpha_t . T_t . op(C_t)   one CTA per (non-empty 32x32 tile of a task, point, re/im)
// The inner product runs on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA) from operand panels staged in shared
// memory (the U,V blocks and the amplitudes of all batched points stay L2-resident).
#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

__device__ __forceinline__ size_t quad_offset(int layout_pack, int c, int k, size_t nxy) {
  // pack layout = Broyden vector order [reX reY imX imY | reP reQ imP imQ] (pnfam_broyden.f90:50-60)
  if (layout_pack) return (k < 2) ? ((size_t)c * 2 + k) * nxy : (4 + (size_t)c * 2 + (k - 2)) * nxy;
  return ((size_t)c * 4 + k) * nxy;  // [re: q0..q3 | im: q0..q3]
}

// Both phases: one CTA (256 threads) per non-empty 32x32 output tile of one point, the host lists the tiles largest
// first.  Warps 0-3 compute the real flavour, warps 4-7 the imaginary one (a 16x16 sub-tile = 2x2 DMMA tiles each): the
// operand that does not depend on the flavour (the U/V block) is staged once for both.  The full-K operand panels are
// staged in shared memory as [k][32 + 4] (zero padded to a multiple of 4 in k, so the DMMA loop has no bounds checks;
// row stride 36 makes the 8-byte fragment loads bank-conflict free) and every staged element feeds 16 DMMA lanes.
constexpr int TR_LD = 36;

__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// panel[k][x] = *src(x0 + x, k), x < 32: `xfast` tells which index is contiguous in memory (coalesced global reads).
// 256 threads.  Every element is an asynchronous 8-byte copy (zero stored directly outside the block), so all loads of
// all panels of a CTA are in flight together; the caller waits with cp_async_wait_all() + __syncthreads().
template <class F>
__device__ __forceinline__ void stage_panel(double* __restrict__ panel, int K4, int K, int X, int x0, bool xfast, F src) {
  const int lo = threadIdx.x & 31, hi = threadIdx.x >> 5;
  if (xfast) {
    const int x = lo;
    const bool xin = x0 + x < X;
    for (int k = hi; k < K4; k += 8) {
      if (xin && k < K) cp_async8(&panel[k * TR_LD + x], src(x0 + x, k));
      else panel[k * TR_LD + x] = 0.0;
    }
  } else {
    for (int k = lo; k < K4; k += 32)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int x = hi + 8 * u;
        if (x0 + x < X && k < K) cp_async8(&panel[k * TR_LD + x], src(x0 + x, k));
        else panel[k * TR_LD + x] = 0.0;
      }
  }
}

// MI x NJ DMMA tiles (8 x 8 each) of a warp's 16 x 16 sub-tile: the tiles that lie wholly outside the block are skipped
template <int MI, int NJ>
__device__ __forceinline__ void panel_gemm_tiles(double (&acc)[2][2][2], const double* __restrict__ ap, const double* __restrict__ bp, int K4) {
#pragma unroll 2
  for (int k0 = 0; k0 < K4; k0 += 4) {
    double a[2], b[2];
#pragma unroll
    for (int i = 0; i < MI; i++) a[i] = ap[k0 * TR_LD + 8 * i];
#pragma unroll
    for (int j = 0; j < NJ; j++) b[j] = bp[k0 * TR_LD + 8 * j];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// mrem / nrem: rows / columns of the block left from the first row / column of the warp's sub-tile (> 0)
__device__ __forceinline__ void panel_gemm_16x16(double (&acc)[2][2][2], const double* __restrict__ As, const double* __restrict__ Bs,
                                                 int K4, int m0, int n0, int mrem, int nrem) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  const double* __restrict__ ap = As + lc * TR_LD + m0 + lr;
  const double* __restrict__ bp = Bs + lc * TR_LD + n0 + lr;
  const bool m2 = mrem > 8, n2 = nrem > 8;
  if (m2 && n2) panel_gemm_tiles<2, 2>(acc, ap, bp, K4);
  else if (m2) panel_gemm_tiles<2, 1>(acc, ap, bp, K4);
  else if (n2) panel_gemm_tiles<1, 2>(acc, ap, bp, K4);
  else panel_gemm_tiles<1, 1>(acc, ap, bp, K4);
}

__device__ __forceinline__ void store_tile(double* __restrict__ O, const double (&v)[2][2][2], int M, int N, int m0, int n0) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int m = m0 + i * 8 + lr, n = n0 + j * 8 + 2 * lc + q;
        if (m < M && n < N) O[m + (size_t)n * M] = v[i][j][q];
      }
}

__global__ void __launch_bounds__(256, 3) transform_phase1_kernel(const DevTask* __restrict__ tasks, const int4* __restrict__ tiles,
                                                              TransformArgs args, int kpad) {
  extern __shared__ __align__(16) double tr_smem[];
  if (args.ctrl && (int)blockIdx.y >= args.ctrl->nactive) return;
  const int4 e = tiles[blockIdx.x];                 // (task, term, first row, first column)
  const DevTask& tk = tasks[e.x];
  const DevTerm& tm = tk.t[e.y];
  const int M = tk.m, N = tk.n, K = M, K4 = (K + 3) & ~3;
  const int za = blockIdx.y, p = args.active[za];
  const int tm0 = e.z, tn0 = e.w;
  const int warp = threadIdx.x >> 5, c = warp >> 2;
  const int m0 = ((warp >> 1) & 1) * 16, n0 = (warp & 1) * 16;
  double* As = tr_smem;                              // op(A): shared by both flavours
  double* Bs = tr_smem + (size_t)(1 + c) * kpad * TR_LD;
  const double* __restrict__ Am = args.W[tm.a_mat] + tm.a_off;
  const int at = tm.a_trans, bt = tm.b_trans;
  stage_panel(As, K4, K, M, tm0, !at, [&](int i, int k) { return at ? Am + k + (size_t)i * M : Am + i + (size_t)k * M; });
#pragma unroll
  for (int cc = 0; cc < 2; cc++) {
    const double* __restrict__ Bm = args.in + (size_t)p * args.in_pstride + quad_offset(args.in_pack, cc, tm.b_quad, args.nxy) + tm.b_off;
    stage_panel(tr_smem + (size_t)(1 + cc) * kpad * TR_LD, K4, K, N, tn0, bt != 0,
                [&](int j, int k) { return bt ? Bm + j + (size_t)k * N : Bm + k + (size_t)j * M; });
  }
  cp_async_wait_all();
  __syncthreads();
  if (tm0 + m0 >= M || tn0 + n0 >= N) return;
  double acc[2][2][2] = {};
  panel_gemm_16x16(acc, As, Bs, K4, m0, n0, M - tm0 - m0, N - tn0 - n0);
  double* __restrict__ T = args.scratch + ((size_t)za * 2 + c) * args.scratch_stride + tm.t_off;
  store_tile(T, acc, M, N, tm0 + m0, tn0 + n0);
}

__global__ void __launch_bounds__(256, 3) transform_phase2_kernel(const DevTask* __restrict__ tasks, const int4* __restrict__ tiles,
                                                              TransformArgs args, int kpad) {
  extern __shared__ __align__(16) double tr_smem[];
  if (args.ctrl && (int)blockIdx.y >= args.ctrl->nactive) return;
  const int4 e = tiles[blockIdx.x];                 // (task, -, first row, first column)
  const DevTask& tk = tasks[e.x];
  const int M = tk.m, N = tk.n, K = N, K4 = (K + 3) & ~3;
  const int za = blockIdx.y, p = args.active[za];
  const int tm0 = e.z, tn0 = e.w;
  const int warp = threadIdx.x >> 5, c = warp >> 2;
  const int m0 = ((warp >> 1) & 1) * 16, n0 = (warp & 1) * 16;
  double* Cs = tr_smem;                              // op(C): shared by both flavours
  double* Ts = tr_smem + (size_t)(1 + c) * kpad * TR_LD;
  const bool live = tm0 + m0 < M && tn0 + n0 < N;
  double out[2][2][2] = {};
  for (int t = 0; t < tk.nterms; t++) {
    const DevTerm& tm = tk.t[t];
    const double* __restrict__ Cm = args.W[tm.c_mat] + tm.c_off;
    const int ct = tm.c_trans;
    if (t > 0) __syncthreads();
    stage_panel(Cs, K4, K, N, tn0, ct != 0, [&](int j, int k) { return ct ? Cm + j + (size_t)k * N : Cm + k + (size_t)j * N; });
#pragma unroll
    for (int cc = 0; cc < 2; cc++) {
      const double* __restrict__ T = args.scratch + ((size_t)za * 2 + cc) * args.scratch_stride + tm.t_off;
      stage_panel(tr_smem + (size_t)(1 + cc) * kpad * TR_LD, K4, K, M, tm0, true, [&](int i, int k) { return T + i + (size_t)k * M; });
    }
    cp_async_wait_all();
    __syncthreads();
    if (live) {
      double acc[2][2][2] = {};
      panel_gemm_16x16(acc, Ts, Cs, K4, m0, n0, M - tm0 - m0, N - tn0 - n0);
      const double alpha = c ? tm.alpha_im : tm.alpha_re;
#pragma unroll
      for (int i = 0; i < 8; i++) (&out[0][0][0])[i] += alpha * (&acc[0][0][0])[i];
    }
  }
  if (!live) return;
  double* __restrict__ O = args.out + (size_t)p * args.out_pstride + quad_offset(args.out_pack, c, tk.out_quad, args.nxy) + tk.out_off;
  store_tile(O, out, M, N, tm0 + m0, tn0 + n0);
}

void launch_transform(const DevicePlan& plan, const TransformArgs& args, int nactive, cudaStream_t stream) {
  if (nactive <= 0) return;
  const int kpad = (plan.max_dim + 3) & ~3;
  const size_t smem = (size_t)3 * kpad * TR_LD * sizeof(double);
  static PerDeviceMax attr;
  if (smem > 48 * 1024 && attr.raise(smem)) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(transform_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(transform_phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 g1(plan.ntiles1, nactive), g2(plan.ntiles2, nactive);
  if (plan.ntiles1 > 0) transform_phase1_kernel<<<g1, 256, smem, stream>>>(plan.tasks, plan.tiles1, args, kpad);
  if (plan.ntiles2 > 0) transform_phase2_kernel<<<g2, 256, smem, stream>>>(plan.tasks, plan.tiles2, args, kpad);
}

}  // namespace pnfam
