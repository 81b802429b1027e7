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


// ================================================================================================
// Fused transform: both products of a triple product in one kernel, the intermediate never leaves the registers.
//
// One CTA (256 threads) = (job, strip of <= 4 DMMA m-tiles = 32 rows of the output block(s), omega point).  Warp w works on
// m-tile (w & 3) of the strip for the real (w < 4) or imaginary (w >= 4) flavour of the middle operand and keeps its
// 8 x N slice of T_x = sum_t alpha_t op(A_t) op(B_t) in registers as DMMA accumulators.  The accumulator fragment of
// mma.m8n8k4 (lane l: row l/4, columns 2(l%4), 2(l%4)+1) IS a valid A fragment of the same instruction for the k-steps
// {0,2,4,6} and {1,3,5,7} of that 8-column tile -- the contraction index may be visited in any order as long as the B
// operand follows -- so phase 2, out_o += beta T_x op(C_{o,x}), runs straight from those registers: no shared-memory or
// global round trip of the intermediate, no shuffles.  op(C) is staged with its rows permuted accordingly.
// op(B) (flavour dependent, k-chunks of 16 rows) and op(C) (shared by the flavours, chunks of 64 rows x 32 columns) stream
// through a 3-stage cp.async ring, one __syncthreads per chunk; op(A) fragments are read straight from global memory
// (each warp needs its own 8 rows once; the U, V blocks are L2 resident) one chunk ahead.  Several groups x are
// accumulated through the output block itself (each thread re-reads only what it wrote).
// ================================================================================================
constexpr int TF_KC = 16;        // rows of a phase-1 chunk (per flavour)
constexpr int TF_K2 = 64;        // contraction rows of a phase-2 chunk
constexpr int TF_PW = 32;        // output columns of a phase-2 panel
constexpr int TF_LD2 = 36;       // row stride of a phase-2 chunk
constexpr int TF_STAGES = 3;
template <int NT> struct TfShape {
  static constexpr int LD = NT * 8 + 4;       // row stride of a phase-1 chunk (= 4 or 12 mod 16: conflict-free fragments)
  static constexpr int CH = (2 * TF_KC * LD > TF_K2 * TF_LD2) ? 2 * TF_KC * LD : TF_K2 * TF_LD2;   // doubles per stage
  static constexpr size_t SMEM = (size_t)TF_STAGES * CH * sizeof(double);
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// DMMA blocks with compile-time tile counts (the tile count of a block is uniform over the CTA: one switch per chunk
// instead of a predicate per DMMA); nks = k-steps of the chunk that hold data (the last chunk of a block may be short)
template <int NJ, int NT, int LD>
__device__ __forceinline__ void tf_phase1_block(double (&T)[NT][2], const double (&a)[4], const double* __restrict__ bp, int nks) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++)
    if (ks < nks) {
#pragma unroll
      for (int j = 0; j < NJ; j++) dmma884(T[j][0], T[j][1], a[ks], bp[ks * 4 * LD + 8 * j]);
    }
}
// Rows of a phase-2 chunk are permuted inside every group of 8 contraction indices so that BOTH access patterns hit 4
// rows that differ modulo 4 (row stride 36 doubles: bank-conflict free): the DMMA fragments -- the accumulator-as-A trick
// visits k = 2 lc + h, lc = 0..3 -- and the 4-consecutive-k staging patches.  k -> row: 0 2 1 3 6 4 7 5.
__device__ __forceinline__ int tf_rowperm(int k) { return (0x57463120 >> (4 * k)) & 7; }
template <int NP>
__device__ __forceinline__ void tf_phase2_tile(double (&acc)[4][2], double t0, double t1, const double* __restrict__ c0, const double* __restrict__ c1) {
#pragma unroll
  for (int jp = 0; jp < NP; jp++) dmma884(acc[jp][0], acc[jp][1], t0, c0[8 * jp]);
#pragma unroll
  for (int jp = 0; jp < NP; jp++) dmma884(acc[jp][0], acc[jp][1], t1, c1[8 * jp]);
}
// the T tiles [KT0, KT0 + 8) of one phase-2 chunk
template <int KT0, int NT>
__device__ __forceinline__ void tf_phase2_block(double (&acc)[4][2], const double (&T)[NT][2], int n8, int np8, const double* __restrict__ q0,
                                                const double* __restrict__ q1) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    constexpr int dummy = 0;
    (void)dummy;
    if (KT0 + i < NT && KT0 + i < n8) {
      constexpr int kt_max = NT - 1;
      const int kt = KT0 + i < NT ? KT0 + i : kt_max;   // (static) keeps the index inside the array for the dead iterations
      const double* __restrict__ c0 = q0 + i * 8 * TF_LD2;
      const double* __restrict__ c1 = q1 + i * 8 * TF_LD2;
      if (np8 == 4) tf_phase2_tile<4>(acc, T[kt][0], T[kt][1], c0, c1);
      else if (np8 == 3) tf_phase2_tile<3>(acc, T[kt][0], T[kt][1], c0, c1);
      else if (np8 == 2) tf_phase2_tile<2>(acc, T[kt][0], T[kt][1], c0, c1);
      else tf_phase2_tile<1>(acc, T[kt][0], T[kt][1], c0, c1);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(256, NT <= 16 ? 2 : 1) transform_fused_kernel(const FusedJob* __restrict__ jobs, const int4* __restrict__ ctas,
                                                                                TransformArgs args) {
  using SH = TfShape<NT>;
  constexpr int LD = SH::LD, CH = SH::CH;
  extern __shared__ __align__(16) double tr_smem[];
  __shared__ FusedJob J;                                // the job descriptor: read many times, from shared memory
  if (args.ctrl && (int)blockIdx.y >= args.ctrl->nactive) return;
  const int4 e = ctas[blockIdx.x];                      // (job, first m-tile, m-tiles of the strip, -)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  {
    const int* src = reinterpret_cast<const int*>(jobs + e.x);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < (int)(sizeof(FusedJob) / 4); i += 256) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int M = J.m, N = J.n, K = M;
  const int za = blockIdx.y, pt = args.active[za];
  const int c = warp >> 2;                              // flavour of this warp
  const int row0 = tf_rowperm(2 * lc), row1 = tf_rowperm(2 * lc + 1);   // rows of this lane's phase-2 fragments
  const bool active = (warp & 3) < e.z;
  const int row = (e.y + (warp & 3)) * 8 + lr;
  const bool row_ok = active && row < M;
  const int nk1 = (K + TF_KC - 1) / TF_KC, nk2 = (N + TF_K2 - 1) / TF_K2, npan = (N + TF_PW - 1) / TF_PW;
  const int n8 = (N + 7) >> 3, ncol = n8 << 3;
  const int ngroups = J.ngroups, nout = J.nout;
  const double* __restrict__ in_pt = args.in + (size_t)pt * args.in_pstride;
  const size_t dflav = quad_offset(args.in_pack, 1, 0, args.nxy) - quad_offset(args.in_pack, 0, 0, args.nxy);   // re -> im
  const double* __restrict__ W0 = args.W[0];
  const double* __restrict__ W1 = args.W[1];
  const double* __restrict__ W2 = args.W[2];
  const double* __restrict__ W3 = args.W[3];
  auto wmat = [&](int m) { return m == 0 ? W0 : (m == 1 ? W1 : (m == 2 ? W2 : W3)); };

  // ---- staging (all threads).  The issue cursor runs TF_STAGES-1 chunks ahead of the math; its position inside the
  //      job is (group, phase, term | output, panel, chunk); the source pointer of the current operand is advanced by
  //      a constant per chunk.  Thread -> element maps (fixed for the whole kernel):
  //        k fastest in memory: a warp copies 4 (k) x 8 (j) patches -- full 32-byte sectors on the global side and the
  //                             bank pattern of the DMMA fragment loads on the shared side (conflict free)
  //        j fastest in memory: a warp copies 32 consecutive j of one row
  const int kq = lane & 3, jq = lane >> 2;
  const int p1_kl = 4 * (warp & 3) + kq, p1_j0 = 8 * (warp >> 2) + jq;           // phase 1, k fastest
  const int p2_kra = 4 * warp + kq, p2_krb = p2_kra + 32;                          // phase 2, k fastest
  const int p2_da = ((p2_kra & ~7) + tf_rowperm(p2_kra & 7)) * TF_LD2 + jq, p2_db = ((p2_krb & ~7) + tf_rowperm(p2_krb & 7)) * TF_LD2 + jq;
  int sx = 0, sph = 0, st = 0, so = 0, sp = 0, sk = 0;   // group, phase (0/1), term, output, panel, chunk
  const double* sptr = nullptr;
  int strans = 0;
  auto begin_operand = [&]() {
    if (sx >= ngroups) return;
    if (sph == 0) {
      const FusedTerm& tm = J.t[sx][st];
      strans = tm.b_trans;
      const double* B0 = in_pt + quad_offset(args.in_pack, 0, tm.b_quad, args.nxy) + tm.b_off;
      sptr = strans ? B0 + lane + (size_t)warp * N : B0 + p1_kl + (size_t)p1_j0 * K;
    } else {
      const FusedOut& fo = J.o[so];
      strans = fo.c_trans[sx];
      const double* Cm = wmat(fo.c_mat[sx]) + fo.c_off[sx];
      const int col0 = sp * TF_PW;
      sptr = strans ? Cm + col0 + lane + (size_t)warp * N : Cm + (size_t)(col0 + jq) * N;
    }
  };
  auto stage = [&](double* __restrict__ buf) {
    if (sph == 0) {                                     // phase 1: rows k0 .. k0+15 of op(B_t), both flavours
      const int k0 = sk * TF_KC;
      if (!strans) {                                    // op(B)[k][j] = B[k + j K]
        double* d = buf + p1_kl * LD + p1_j0;
        int j = p1_j0;
        if (k0 + p1_kl < K) {
          const double* s0 = sptr;
          for (; j < N; j += 16, d += 16, s0 += 16 * (size_t)K) { cp_async8(d, s0); cp_async8(d + TF_KC * LD, s0 + dflav); }
        }
        for (; j < ncol; j += 16, d += 16) { d[0] = 0.0; d[TF_KC * LD] = 0.0; }
        sptr += TF_KC;
      } else {                                          // op(B)[k][j] = B[j + k N]
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int kr = warp + 8 * h;
          double* d = buf + kr * LD + lane;
          int j = lane;
          if (k0 + kr < K) {
            const double* s0 = sptr + (size_t)(8 * h) * N;
            for (; j < N; j += 32, d += 32, s0 += 32) { cp_async8(d, s0); cp_async8(d + TF_KC * LD, s0 + dflav); }
          }
          for (; j < ncol; j += 32, d += 32) { d[0] = 0.0; d[TF_KC * LD] = 0.0; }
        }
        sptr += (size_t)TF_KC * N;
      }
    } else {                                            // phase 2: 64 rows x 32 columns of op(C_{o,x}), rows permuted
      const int k0 = sk * TF_K2, col0 = sp * TF_PW;
      if (!strans) {                                    // op(C)[k][j] = C[k + j N]
        const bool ka = k0 + p2_kra < N, kb = k0 + p2_krb < N;
        const double* s0 = sptr;
#pragma unroll
        for (int u = 0; u < 4; u++, s0 += 8 * (size_t)N) {
          const bool jin = col0 + 8 * u + jq < N;
          if (ka && jin) cp_async8(buf + p2_da + 8 * u, s0 + p2_kra); else buf[p2_da + 8 * u] = 0.0;
          if (kb && jin) cp_async8(buf + p2_db + 8 * u, s0 + p2_krb); else buf[p2_db + 8 * u] = 0.0;
        }
        sptr += TF_K2;
      } else {                                          // op(C)[k][j] = C[j + k N]
        const bool jin = col0 + lane < N;
        const double* s0 = sptr;
#pragma unroll
        for (int u = 0; u < 8; u++, s0 += 8 * (size_t)N) {
          const int kr = warp + 8 * u;                  // rows 8u + warp: the permutation depends on the warp only
          double* d = buf + (8 * u + tf_rowperm(warp)) * TF_LD2 + lane;
          if (jin && k0 + kr < N) cp_async8(d, s0);
          else d[0] = 0.0;
        }
        sptr += (size_t)TF_K2 * N;
      }
    }
    // advance the cursor
    sk++;
    if (sph == 0) {
      if (sk == nk1) { sk = 0; if (++st == J.nterms[sx]) { st = 0; sph = 1; so = 0; sp = 0; } begin_operand(); }
    } else if (sk == nk2) {
      sk = 0;
      if (++sp == npan) { sp = 0; if (++so == nout) { so = 0; sph = 0; sx++; } }
      begin_operand();
    }
  };
  begin_operand();

  // ---- A fragments of one chunk: op(A)[row][k0 + 4 ks + lc], fetched one chunk ahead, scaled by alpha at hand-over ----
  const double* ap = nullptr;       // this lane's element of chunk 0
  size_t astep = 0;                 // distance of two k
  auto begin_a = [&](const FusedTerm& tm) {
    const double* Am = wmat(tm.a_mat) + tm.a_off;
    astep = tm.a_trans ? 1 : (size_t)M;
    ap = tm.a_trans ? Am + lc + (size_t)row * M : Am + row + (size_t)lc * M;
  };
  auto load_a = [&](int k0, double (&a)[4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const int k = k0 + ks * 4;
      a[ks] = (row_ok && k + lc < K) ? __ldg(ap + (size_t)k * astep) : 0.0;
    }
  };

  // ---- the math walks the same chunk sequence (one call site of the staging code) --------------------------
  int total = 0;
  for (int x = 0; x < ngroups; x++) total += J.nterms[x] * nk1 + nout * npan * nk2;
  int cx = 0, cph = 0, ct = 0, co = 0, cp_ = 0, ck = 0;  // math cursor: group, phase, term, output, panel, chunk
  double T[NT][2], acc[4][2], a_cur[4], a_nxt[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < NT; j++) T[j][0] = T[j][1] = 0.0;
#pragma unroll
  for (int jp = 0; jp < 4; jp++) acc[jp][0] = acc[jp][1] = 0.0;
  double alpha_nxt;
  {
    const FusedTerm& tm = J.t[0][0];
    begin_a(tm);
    load_a(0, a_cur);
    const double al = c ? tm.alpha_im : tm.alpha_re;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) a_cur[ks] *= al;
    alpha_nxt = al;
  }
  int ibuf = 0, cbuf = 0;
  for (int ch = -(TF_STAGES - 1); ch < total; ch++) {
    if (ch >= 0) {
      cp_async_wait_group<TF_STAGES - 2>();             // chunk `ch` has landed for this thread ...
      __syncthreads();                                  // ... and for all; the buffer of chunk ch-1 is free again
    }
    if (sx < ngroups) stage(tr_smem + (size_t)ibuf * CH);
    cp_async_commit();
    if (++ibuf == TF_STAGES) ibuf = 0;
    if (ch < 0) continue;
    const double* __restrict__ buf = tr_smem + (size_t)cbuf * CH;
    if (++cbuf == TF_STAGES) cbuf = 0;
    if (cph == 0) {
      // ---- phase 1: T += alpha op(A_t)[rows][k0..] op(B_t)[k0..][:] ------------------------------------------
      const bool more_k = ck + 1 < nk1, more_t = ct + 1 < J.nterms[cx];
      if (more_k) load_a((ck + 1) * TF_KC, a_nxt);
      else if (more_t) { const FusedTerm& tm = J.t[cx][ct + 1]; begin_a(tm); load_a(0, a_nxt); alpha_nxt = c ? tm.alpha_im : tm.alpha_re; }
      if (active) {
        const double* __restrict__ bp = buf + (size_t)c * TF_KC * LD + lc * LD + lr;
        const int nks = min(4, (K - ck * TF_KC + 3) >> 2);
        switch (n8) {
#define TF_CASE(n) case n: if (n <= NT) tf_phase1_block<(n <= NT ? n : 1), NT, LD>(T, a_cur, bp, nks); break;
          TF_CASE(1) TF_CASE(2) TF_CASE(3) TF_CASE(4) TF_CASE(5) TF_CASE(6) TF_CASE(7) TF_CASE(8) TF_CASE(9) TF_CASE(10) TF_CASE(11)
          TF_CASE(12) TF_CASE(13) TF_CASE(14) TF_CASE(15) TF_CASE(16) TF_CASE(17) TF_CASE(18) TF_CASE(19) TF_CASE(20) TF_CASE(21) TF_CASE(22)
#undef TF_CASE
          default: break;
        }
      }
#pragma unroll
      for (int ks = 0; ks < 4; ks++) a_cur[ks] = alpha_nxt * a_nxt[ks];
      if (++ck == nk1) { ck = 0; if (++ct == J.nterms[cx]) { ct = 0; cph = 1; co = 0; cp_ = 0; } }
    } else {
      // ---- phase 2: acc += T[:, k0..] op(C_{o,x})[k0..][panel] ----------------------------------------------
      const int np8 = min(4, n8 - 4 * cp_);             // n-tiles of this panel
      if (active) {
        const double* __restrict__ q0 = buf + row0 * TF_LD2 + lr;
        const double* __restrict__ q1 = buf + row1 * TF_LD2 + lr;
        if (ck == 0) tf_phase2_block<0, NT>(acc, T, n8, np8, q0, q1);
        else if (ck == 1) tf_phase2_block<8, NT>(acc, T, n8, np8, q0, q1);
        else tf_phase2_block<16, NT>(acc, T, n8, np8, q0, q1);
      }
      if (++ck == nk2) {
        ck = 0;
        // panel finished: out (+)= beta acc
        const FusedOut& fo = J.o[co];
        if (row_ok) {
          const double beta = c ? fo.beta_im[cx] : fo.beta_re[cx];
          const int col0 = cp_ * TF_PW;
          double* __restrict__ dst = args.out + (size_t)pt * args.out_pstride + quad_offset(args.out_pack, c, fo.out_quad, args.nxy) + fo.out_off +
                                     row + (size_t)(col0 + 2 * lc) * M;
#pragma unroll
          for (int jp = 0; jp < 4; jp++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
              if (col0 + 8 * jp + 2 * lc + q < N) {
                double* __restrict__ d2 = dst + (size_t)(8 * jp + q) * M;
                double v = beta * acc[jp][q];
                if (cx > 0) v += *d2;
                *d2 = v;
              }
            }
        }
#pragma unroll
        for (int jp = 0; jp < 4; jp++) acc[jp][0] = acc[jp][1] = 0.0;
        if (++cp_ == npan) {
          cp_ = 0;
          if (++co == nout) {                           // group finished: next group starts with a fresh T
            co = 0; cph = 0; cx++;
#pragma unroll
            for (int j = 0; j < NT; j++) T[j][0] = T[j][1] = 0.0;
            if (cx < ngroups) {
              const FusedTerm& tm = J.t[cx][0];
              begin_a(tm);
              load_a(0, a_cur);
              const double al = c ? tm.alpha_im : tm.alpha_re;
#pragma unroll
              for (int ks = 0; ks < 4; ks++) a_cur[ks] *= al;
              alpha_nxt = al;
            }
          }
        }
      }
    }
  }
  cp_async_wait_all();
}

template <int NT>
static void launch_fused(const DevicePlan& plan, int cls, const TransformArgs& args, int nactive, cudaStream_t stream) {
  static PerDeviceMax attr;
  const size_t smem = TfShape<NT>::SMEM;
  if (attr.raise(smem))
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(transform_fused_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  transform_fused_kernel<NT><<<dim3(plan.nfctas[cls], nactive), 256, smem, stream>>>(plan.jobs, plan.fctas[cls], args);
}

int launch_transform(const DevicePlan& plan, const TransformArgs& args, int nactive, cudaStream_t stream) {
  if (nactive <= 0) return 0;
  if (plan.fused) {
    int n = 0;
    if (plan.nfctas[2] > 0) { launch_fused<22>(plan, 2, args, nactive, stream); n++; }
    if (plan.nfctas[1] > 0) { launch_fused<16>(plan, 1, args, nactive, stream); n++; }
    if (plan.nfctas[0] > 0) { launch_fused<11>(plan, 0, args, nactive, stream); n++; }
    return n;
  }
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
  return (plan.ntiles1 > 0) + (plan.ntiles2 > 0);
}

}  // namespace pnfam
