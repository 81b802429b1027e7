// Blocks (a) and (d) of the FAM iteration: quasiparticle <-> single-particle transforms
//     dRsp = W_a . dRqp . W_b^T       (reference: pnfam_solver.f90:144-150 -> triprod_bbm -> 2 dgemm per block,
//     dHqp = W_a^T . dHsp . W_b        pnfam_solver.f90:160-166,       pnfam_type_blockmatrix.f90:202-206)
// executed as two grouped-GEMM launches over the task list built by host/symbolic.cpp:
//   phase 1:  T_t   = op(A_t) . op(B_t)               one CTA tile per (task term, 32x32 tile, point, re/im)
//   phase 2:  out   = sum_t alpha_t . T_t . op(C_t)   one CTA tile per (task, 32x32 tile, point, re/im)
// The inner product runs on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA); operands are read
// through L1/L2 (the U,V blocks and the amplitudes of all batched points stay L2-resident).
#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

__device__ __forceinline__ size_t quad_offset(int layout_pack, int c, int k, size_t nxy) {
  // pack layout = Broyden vector order [reX reY imX imY | reP reQ imP imQ] (pnfam_broyden.f90:50-60)
  if (layout_pack) return (k < 2) ? ((size_t)c * 2 + k) * nxy : (4 + (size_t)c * 2 + (k - 2)) * nxy;
  return ((size_t)c * 4 + k) * nxy;  // [re: q0..q3 | im: q0..q3]
}

// One warp computes a 16x16 sub-tile (2x2 DMMA tiles) of C = A(MxK) * B(KxN) with generic accessors.
template <class FA, class FB>
__device__ __forceinline__ void warp_gemm_16x16(double (&acc)[2][2][2], int m0, int n0, int M, int N, int K, FA A, FB B) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  for (int k0 = 0; k0 < K; k0 += 4) {
    const int k = k0 + lc;
    double a[2], b[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int m = m0 + i * 8 + lr;
      a[i] = (m < M && k < K) ? A(m, k) : 0.0;
      const int n = n0 + i * 8 + lr;
      b[i] = (n < N && k < K) ? B(k, n) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

struct Phase1Entry {
  int task, term;
};

__global__ void __launch_bounds__(128) transform_phase1_kernel(const DevTask* __restrict__ tasks, const Phase1Entry* __restrict__ entries,
                                                              TransformArgs args) {
  const Phase1Entry e = entries[blockIdx.y];
  const DevTask& tk = tasks[e.task];
  const DevTerm& tm = tk.t[e.term];
  const int M = tk.m, N = tk.n;
  const int tiles_n = (N + 31) / 32, tiles_m = (M + 31) / 32;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int z = blockIdx.z, c = z & 1, p = args.active[z >> 1];
  const int tm0 = (blockIdx.x / tiles_n) * 32, tn0 = (blockIdx.x % tiles_n) * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = tm0 + (warp >> 1) * 16, n0 = tn0 + (warp & 1) * 16;
  const double* __restrict__ Am = args.W[tm.a_mat] + tm.a_off;
  const double* __restrict__ Bm = args.in + (size_t)p * args.in_pstride + quad_offset(args.in_pack, c, tm.b_quad, args.nxy) + tm.b_off;
  double acc[2][2][2] = {};
  const int at = tm.a_trans, bt = tm.b_trans;
  warp_gemm_16x16(
      acc, m0, n0, M, N, M,
      [&](int i, int k) { return at ? Am[k + (size_t)i * M] : Am[i + (size_t)k * M]; },
      [&](int k, int j) { return bt ? Bm[j + (size_t)k * N] : Bm[k + (size_t)j * M]; });
  double* __restrict__ T = args.scratch + ((size_t)(z >> 1) * 2 + c) * args.scratch_stride + tm.t_off;
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int m = m0 + i * 8 + lr, n = n0 + j * 8 + 2 * lc + q;
        if (m < M && n < N) T[m + (size_t)n * M] = acc[i][j][q];
      }
}

__global__ void __launch_bounds__(128) transform_phase2_kernel(const DevTask* __restrict__ tasks, TransformArgs args) {
  const DevTask& tk = tasks[blockIdx.y];
  const int M = tk.m, N = tk.n;
  const int tiles_n = (N + 31) / 32, tiles_m = (M + 31) / 32;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int z = blockIdx.z, c = z & 1, p = args.active[z >> 1];
  const int tm0 = (blockIdx.x / tiles_n) * 32, tn0 = (blockIdx.x % tiles_n) * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = tm0 + (warp >> 1) * 16, n0 = tn0 + (warp & 1) * 16;
  double out[2][2][2] = {};
  for (int t = 0; t < tk.nterms; t++) {
    const DevTerm& tm = tk.t[t];
    const double* __restrict__ T = args.scratch + ((size_t)(z >> 1) * 2 + c) * args.scratch_stride + tm.t_off;
    const double* __restrict__ Cm = args.W[tm.c_mat] + tm.c_off;
    const int ct = tm.c_trans;
    double acc[2][2][2] = {};
    warp_gemm_16x16(
        acc, m0, n0, M, N, N, [&](int i, int k) { return T[i + (size_t)k * M]; },
        [&](int k, int j) { return ct ? Cm[j + (size_t)k * N] : Cm[k + (size_t)j * N]; });
    const double alpha = c ? tm.alpha_im : tm.alpha_re;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        out[i][j][0] += alpha * acc[i][j][0];
        out[i][j][1] += alpha * acc[i][j][1];
      }
  }
  double* __restrict__ O = args.out + (size_t)p * args.out_pstride + quad_offset(args.out_pack, c, tk.out_quad, args.nxy) + tk.out_off;
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int m = m0 + i * 8 + lr, n = n0 + j * 8 + 2 * lc + q;
        if (m < M && n < N) O[m + (size_t)n * M] = out[i][j][q];
      }
}

void launch_transform(const DevicePlan& plan, const TransformArgs& args, int nactive, cudaStream_t stream) {
  if (nactive <= 0) return;
  const int tiles = ((plan.max_dim + 31) / 32) * ((plan.max_dim + 31) / 32);
  dim3 g1(tiles, plan.nentries, nactive * 2), g2(tiles, plan.ntasks, nactive * 2);
  transform_phase1_kernel<<<g1, 128, 0, stream>>>(plan.tasks, reinterpret_cast<const Phase1Entry*>(plan.entries), args);
  transform_phase2_kernel<<<g2, 128, 0, stream>>>(plan.tasks, args);
}

}  // namespace pnfam
