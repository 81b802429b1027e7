// Sum-factorised density and projection kernels (blocks (b) and (c) of the FAM iteration) for a separable basis.
// Same results as density / meanfield / pairingfield of the reference (exes/pnfam/pnfam_hamiltonian_blas.f90:124-711,
// 717-1169, 1175-1262) and as the general-table kernels of hamiltonian.cu, different operation order.
//
// In the harmonic-oscillator basis every table is a product, phi^t_a(ih, il) = Z^t(n_z(a), ih) R^t_a(il)
// (HFBTHO builds them that way, hfbtho_solver.f90:3463-3671), with at most ~N_sh/2 distinct n_z inside one
// (block, spin) segment.  The reference (and hamiltonian.cu) contract  sum_a phi_a(r) rho_ab  over all states a of a
// segment for every grid point: 2 Ng d_a d_b flops per type.  Here, for one Gauss-Laguerre node il at a time:
//
//   density     T^j[k][b]    = sum_{a in slot k} R^j_a(il) rho_ab               FMA,  3 d_a d_b       (j: R0, R1, R2)
//               A^t(ih, b)   = sum_k Z^t(k, ih) T^j(t)[k][b]                    DMMA, M = ih, K = #n_z slots, N = (b, re/im)
//               D^{tt'}(ih) += A^t(ih, b) Z^t'(n_z(b), ih) R^t'_b(il)           FMA epilogue on the accumulator fragments
//   projection  G^t(ih, b)   = sum_t' mf^{tt'}(ih, il) Z^t'(n_z(b), ih) R^t'_b  FMA
//               W^w[k][b]    = sum_ih Z(k, ih) G(ih, b)                         DMMA, M = n_z slots, K = ih, N = (b, re/im)
//               h_ab        += sum_w R^w_a(il) W^w[slot(a)][b]                  FMA,  4 d_a d_b
//
// which executes 2.1x (projection) to 2.4x (density) fewer pipe cycles at 16 shells and moves no Ng x N table at all:
// a CTA reads the rho images (density) or the field tensor of its il (projection) plus a few KB of factors.
// FP64 FMA and DMMA share the issue port of an SM sub-partition, so neither kernel can hide one
// behind the other inside a warp; instead the density is warp-specialised (copy issuer / T-phase producers / DMMA
// consumers, mbarrier hand-over) and the projection runs eight independent warp pairs per CTA whose phases drift
// against each other.  Operand copies are cp.async.bulk completing on mbarriers.  The kernels are instantiated with
// compile-time strides for the common 40-point Gauss-Hermite grid (the DMMA loops are instruction bound: immediates
// instead of integer multiplies are worth 8 %), the kappa / Delta passes (a quarter of the work of rho / h) have their
// own warp-role split and four Gauss-Laguerre nodes per iteration, and the independent launches of a stage go to
// side streams so that they fill the SMs together when few omega points are still active.
#include <algorithm>
#include <vector>

#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

__host__ __device__ constexpr bool sf_mf_nonzero(int t, int t2) { return t == 0 || t2 == 0 || (t < 4 && t2 < 4); }
__host__ __device__ constexpr int sf_mf_pair(int t, int t2) { return t == 0 ? t2 : (t < 4 ? 5 + (t - 1) * 4 + t2 : 17); }

// ================================================================================================
// packed rho / kappa images: image(step)[a][2 b + c], rows a of the segment in slot order, padded columns zero
// ================================================================================================
__global__ void __launch_bounds__(256) sf_pack_kernel(HamArgs g) {
  const SfDev& S = g.sf;
  const int list = blockIdx.y, kind = list >> 1, q = list & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;   // the host sizes the grid with a stale upper bound of the active slots
  if ((int)blockIdx.x >= S.nsteps[list]) return;
  const SfDensStep d = S.steps[list][blockIdx.x];
  const int p = g.active[za];
  const int quad = kind ? g.kap_quad[q] : g.rho_quad[q];
  const double* __restrict__ src0 = g.rsp + ((size_t)p * 2 + 0) * 4 * g.nxy + (size_t)quad * g.nxy + d.rho_off;
  const double* __restrict__ src1 = g.rsp + ((size_t)p * 2 + 1) * 4 * g.nxy + (size_t)quad * g.nxy + d.rho_off;
  double* __restrict__ dst = S.pk[kind] + ((size_t)za * 2 + q) * S.pk_stride[kind] + d.img_off;
  const int ncol = 2 * d.nbc, total = d.na * ncol;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int a = idx / ncol, col = idx - a * ncol;
    const int la = S.p2l[d.a_row0 + a], lb = S.p2l[d.b_row0 + (col >> 1)];
    dst[idx] = lb >= 0 ? ((col & 1) ? src1 : src0)[(size_t)la + (size_t)lb * d.ld] : 0.0;
  }
}

// ================================================================================================
// density
// ================================================================================================
constexpr int SF_DSTAGES = 4;   // operand stages of the density kernel (as many as fit, at least 2)
constexpr int SF_TSTAGES = 2;   // T stages (at most; SfDensLayout::ntst in use): the producers may run this many steps ahead
                                // of the DMMA warps (measured: 3 T stages + 3 operand stages is 2 % slower than 2 + 4)
constexpr int SF_DPROD = 3;     // producer warps (phase T) of the rho pass; the kappa pass, whose DMMA work is a quarter, runs
constexpr int SF_DCONS = 12;    // 5 producer + 10 consumer warps when its m-tiles fit (measured: its consumers waited 31 %
                                // of their time for T with 3 producers)
struct SfDensLayout {
  int off_stage[SF_DSTAGES], off_T[SF_TSTAGES], off_bar, off_x;   // byte offsets in dynamic shared memory
  int nst, ntst;                         // operand stages / T stages in use
  int st_zb, st_ra, st_rb, st_rho;       // inside an operand stage: segtab at 0
  int t_rb, t_zb, t_sz;                  // inside a T stage: T at 0
  int ts;                                // row stride of T (doubles), ts % 16 == 4
  int total;
};

static SfDensLayout make_dens_layout(const SfDev& S) {
  SfDensLayout L{};
  auto up = [](int x) { return (x + 127) & ~127; };
  L.ts = 2 * S.nbc_max + 4;
  L.st_zb = SF_SEGTAB * 4;
  L.st_ra = L.st_zb + S.nbc_max * 4;
  L.st_rb = L.st_ra + S.na_max * 64;
  L.st_rho = L.st_rb + S.nbc_max * 64;
  const int stage = up(L.st_rho + S.na_max * 2 * S.nbc_max * 8);
  L.t_rb = 2 * 3 * S.kpad_max * L.ts * 8;
  L.t_zb = L.t_rb + S.nbc_max * 64;
  L.t_sz = L.t_zb + S.nbc_max * 4;
  const int tstage = up(L.t_sz + SF_KMAX * 4);
  // SF_TSTAGES T stages when at least three operand stages still fit next to them, else two
  for (L.ntst = SF_TSTAGES; L.ntst >= 2; L.ntst--) {
    int off = up(2 * S.nzrows * S.zs * 8);
    for (int i = 0; i < L.ntst; i++) { L.off_T[i] = off; off += tstage; }
    L.off_bar = off; off += 128;
    L.off_x = off; off += 6 * 8 * 32 * 8;          // exchange buffer of the m-tiles shared by two warps
    const int room = 227 * 1024 - off;
    L.nst = std::max(2, std::min(SF_DSTAGES, room / stage));
    for (int i = 0; i < L.nst; i++) { L.off_stage[i] = off; off += stage; }
    L.total = off;
    if (L.ntst == 2 || room / stage >= 3) break;
  }
  return L;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// MODE 0: rho -> D^{tt'}_{ss'} (4 x 4 derivative types);  MODE 1: kappa -> K_{ss'} (plain wave functions).
// One CTA = two Gauss-Laguerre nodes (il pair) of one (pass, omega); it walks the step list of the pass.
// Warp roles (no CTA-wide barrier in the loop, hand-over by mbarriers):
//   warp 15       issues the operand copies of a step (7 linear bulk copies) up to nst steps ahead
//   warps 12..14  phase T: T^j[il][slot][col] = sum_{a in slot} R^j_a(il) rho[a][col]  (FP64 FMA), SF_TSTAGES buffers
//   warps 0..11   DMMA + epilogue of one m-tile (8 grid points) each; with 2 mt = 10 m-tiles the last two are shared
//                 by two warps (half of the n-tiles each) so that every SM sub-partition carries the same DMMA load
// TS_, KPAD_, ZS_ (0 = take the run-time value): row stride of T, padded slot count and row stride of the z tables as
// compile-time constants for the common shapes, so that the fragment addresses of the DMMA loop are immediate offsets
// (the kernel sits at the 128-register limit and would otherwise recompute them with integer multiplies per load).
template <int MODE, int TS_, int KPAD_, int ZS_>
__global__ void __launch_bounds__(SF_THREADS, 1) sf_density_kernel(HamArgs g, SfDensLayout L) {
  constexpr int NJ = MODE == 0 ? 3 : 1;   // radial factor types entering T
  constexpr int NT = MODE == 0 ? 4 : 1;   // derivative types
  extern __shared__ __align__(128) unsigned char smem[];
  const SfDev& S = g.sf;
  const int ilp = blockIdx.x, q = blockIdx.y, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;   // the host sizes the grid with a stale upper bound of the active slots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int list = MODE * 2 + q;
  const SfDensStep* __restrict__ steps = S.steps[list];
  const int nsteps = S.nsteps[list];
  const double* __restrict__ pk = S.pk[MODE] + ((size_t)za * 2 + q) * S.pk_stride[MODE];
  double* Zs = reinterpret_cast<double*>(smem);                 // [2][nzrows][zs]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + L.off_bar);
  unsigned long long *st_full = bars, *st_empty = bars + SF_DSTAGES, *t_full = bars + 2 * SF_DSTAGES, *t_empty = bars + 2 * SF_DSTAGES + SF_TSTAGES;
  const int zs = ZS_ ? ZS_ : S.zs, nzr = S.nzrows, kpad_max = KPAD_ ? KPAD_ : S.kpad_max, ts = TS_ ? TS_ : L.ts, nst = L.nst, ntst = L.ntst;
  const int il0 = 2 * ilp, il1 = min(il0 + 1, S.ngl - 1);
  // consumer roles
  const int M = 2 * S.mt;                                       // m-tiles of the il pair (<= 12)
  const int dcons = (MODE == 1 && M >= 6 && M <= 10) ? 10 : SF_DCONS, dprod = 15 - dcons;   // warp roles (warp 15: copies)
  const int ncons = M >= 6 ? dcons : 2 * M;
  const int first_split = M >= 6 ? 2 * M - dcons : 0;           // m-tiles >= first_split are shared by two warps

  for (int i = tid; i < 2 * nzr * zs; i += SF_THREADS) Zs[i] = S.zt[i];
  if (tid == 0) {
    for (int i = 0; i < SF_DSTAGES; i++) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], dprod); }
    for (int i = 0; i < SF_TSTAGES; i++) { mbar_init(&t_full[i], dprod); mbar_init(&t_empty[i], ncons); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == 15) {
    // ---- operand movement: segment table, z rows of the columns, radial factors of the a rows and of the b columns
    //      for both il, the packed rho image
    //      The step descriptor of k+1 is fetched while step k is issued, and the five copies of a step are issued by
    //      five lanes side by side (a bulk copy costs ~100 clk of issue time in its thread).
    SfDensStep dn = nsteps > 0 ? steps[0] : SfDensStep{};
    int sg = 0, sph = 0;
    for (int k = 0; k < nsteps; k++) {
      const SfDensStep d = dn;
      if (k + 1 < nsteps) dn = steps[k + 1];
      if (k >= nst) mbar_wait(&st_empty[sg], sph ^ 1);       // stage sg, ring pass parity sph (no division per step)
      unsigned char* st = smem + L.off_stage[sg];
      unsigned long long* bar = &st_full[sg];
      const unsigned ra_b = (unsigned)d.na * 64, rb_b = (unsigned)d.nbc * 64, rho_b = (unsigned)d.na * d.nbc * 16;
      if (lane == 0) mbar_expect_tx(bar, SF_SEGTAB * 4 + (unsigned)d.nbc * 4 + ra_b + rb_b + rho_b);
      __syncwarp();
      {
        // one instruction, per-lane operands: lane j moves operand j
        const void* src = pk + d.img_off;
        unsigned char* dst = st + L.st_rho;
        unsigned bytes = rho_b;
        if (lane == 1) { src = S.segtab + (size_t)d.seg_a * SF_SEGTAB; dst = st; bytes = SF_SEGTAB * 4; }
        if (lane == 2) { src = S.zrow + d.b_row0; dst = st + L.st_zb; bytes = (unsigned)d.nbc * 4; }
        if (lane == 3) { src = S.rgp + ((size_t)ilp * S.dqp_p + d.a_row0) * 8; dst = st + L.st_ra; bytes = ra_b; }
        if (lane == 4) { src = S.rgp + ((size_t)ilp * S.dqp_p + d.b_row0) * 8; dst = st + L.st_rb; bytes = rb_b; }
        if (lane < 5) bulk_g2s(dst, src, bytes, bar);
      }
      if (++sg == nst) { sg = 0; sph ^= 1; }
    }
    return;
  }

  if (warp >= dcons) {
    // ---- phase T (producers); forwards what the consumers need from the operand stage, which is recycled earlier
    const int ptid = tid - dcons * 32, pw = warp - dcons;
    const int NP = dprod * 32;
    int sg = 0, sph = 0, tg = 0, tph = 0;                   // stage / ring pass parity of the operand and T rings
    SfDensStep dn = nsteps > 0 ? steps[0] : SfDensStep{};
    for (int k = 0; k < nsteps; k++) {
      const SfDensStep d = dn;
      if (k + 1 < nsteps) dn = steps[k + 1];
      mbar_wait(&st_full[sg], sph);
      if (k >= ntst) mbar_wait(&t_empty[tg], tph ^ 1);
      const unsigned char* st = smem + L.off_stage[sg];
      unsigned char* tst = smem + L.off_T[tg];
      const int* __restrict__ segt = reinterpret_cast<const int*>(st);
      const double* __restrict__ Ra = reinterpret_cast<const double*>(st + L.st_ra);      // [a][il][4]
      const double2* __restrict__ rho = reinterpret_cast<const double2*>(st + L.st_rho);  // [a][b] (re, im)
      double* __restrict__ T = reinterpret_cast<double*>(tst);
      const int kpad = (d.nslots + 3) & ~3;
      // a lane owns one column b (re, im); a warp owns every third n_z slot: the radial factors are warp-uniform
      if (lane < d.nbc) {
        for (int slot = pw; slot < kpad; slot += dprod) {
          int a0 = 0, a1 = 0;
          if (slot < d.nslots) { a0 = segt[slot]; a1 = segt[slot + 1]; }
          double acc[2][NJ][2];
#pragma unroll
          for (int i = 0; i < 4 * NJ; i++) (&acc[0][0][0])[i] = 0.0;
#pragma unroll 2
          for (int a = a0; a < a1; a++) {
            const double2 v = rho[a * d.nbc + lane];
#pragma unroll
            for (int il = 0; il < 2; il++)
#pragma unroll
              for (int j = 0; j < NJ; j++) {
                const double r = Ra[a * 8 + il * 4 + j];
                acc[il][j][0] += r * v.x; acc[il][j][1] += r * v.y;
              }
          }
#pragma unroll
          for (int il = 0; il < 2; il++)
#pragma unroll
            for (int j = 0; j < NJ; j++)
              *reinterpret_cast<double2*>(&T[((il * 3 + j) * kpad_max + slot) * ts + 2 * lane]) = make_double2(acc[il][j][0], acc[il][j][1]);
        }
      }
      const double* __restrict__ Rb = reinterpret_cast<const double*>(st + L.st_rb);
      double* __restrict__ Rbt = reinterpret_cast<double*>(tst + L.t_rb);
      for (int t = ptid; t < d.nbc * 8; t += NP) Rbt[t] = Rb[t];
      if (ptid < d.nbc) reinterpret_cast<int*>(tst + L.t_zb)[ptid] = reinterpret_cast<const int*>(st + L.st_zb)[ptid];
      if (ptid >= 64 && ptid < 64 + SF_KMAX) reinterpret_cast<int*>(tst + L.t_sz)[ptid - 64] = segt[17 + ptid - 64];
      __syncwarp();
      if (lane == 0) { mbar_arrive(&t_full[tg]); mbar_arrive(&st_empty[sg]); }
      if (++sg == nst) { sg = 0; sph ^= 1; }
      if (++tg == ntst) { tg = 0; tph ^= 1; }
    }
    return;
  }

  // ---- consumers: warp -> (m-tile, half of the n-tiles)
  int m, half = 0;
  bool split;
  if (M >= 6) {
    if (warp < M) { m = warp; split = m >= first_split; }
    else { m = M - 1 - (warp - M); split = true; half = 1; }
  } else {
    if (warp >= 2 * M) return;
    m = warp < M ? warp : warp - M; split = true; half = warp < M ? 0 : 1;
  }
  const int ilc = m / S.mt, ih = (m % S.mt) * 8 + lr;
  double acc[NT][NT][2];
#pragma unroll
  for (int i = 0; i < NT * NT * 2; i++) (&acc[0][0][0])[i] = 0.0;
  SfDensStep dn = nsteps > 0 ? steps[0] : SfDensStep{};
  int tg = 0, tph = 0;
  for (int k = 0; k < nsteps; k++) {
    const SfDensStep d = dn;
    if (k + 1 < nsteps) dn = steps[k + 1];
    mbar_wait(&t_full[tg], tph);
    const unsigned char* tst = smem + L.off_T[tg];
    const double* __restrict__ T = reinterpret_cast<const double*>(tst) + (size_t)ilc * 3 * kpad_max * ts;
    const double* __restrict__ Rb = reinterpret_cast<const double*>(tst + L.t_rb) + ilc * 4;   // [b][il][4]
    const int* __restrict__ zb = reinterpret_cast<const int*>(tst + L.t_zb);
    const int* __restrict__ sz = reinterpret_cast<const int*>(tst + L.t_sz);
    const int ksteps = (d.nslots + 3) >> 2, ntn = d.nbc >> 2;
    const int nt_mid = (ntn + 1) >> 1;
    const int nt0 = split ? (half ? nt_mid : 0) : 0, nt1 = split ? (half ? ntn : nt_mid) : ntn;
    constexpr int KS = (KPAD_ ? KPAD_ : SF_KMAX) / 4;       // k-steps a step can have
    double a0[KS], a1[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      a0[ks] = 0.0; a1[ks] = 0.0;
      if (ks < ksteps) {
        const int slot = ks * 4 + lc;
        const int zr = slot < d.nslots ? sz[slot] : 0;
        a0[ks] = Zs[zr * zs + ih];
        if (MODE == 0) a1[ks] = Zs[(nzr + zr) * zs + ih];
      }
    }
    const size_t tj = (size_t)kpad_max * ts;
    for (int nt = nt0; nt < nt1; nt++) {
      double C[NT][2];
#pragma unroll
      for (int i = 0; i < NT * 2; i++) (&C[0][0])[i] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        if (ks < ksteps) {
          const double* __restrict__ tb = T + (size_t)(ks * 4 + lc) * ts + nt * 8 + lr;
          const double b0 = tb[0];
          dmma884(C[0][0], C[0][1], a0[ks], b0);
          if (MODE == 0) {
            const double b1 = tb[tj], b2 = tb[2 * tj];
            dmma884(C[1][0], C[1][1], a0[ks], b1);
            dmma884(C[2][0], C[2][1], a0[ks], b2);
            dmma884(C[3][0], C[3][1], a1[ks], b0);
          }
        }
      }
      // epilogue: contract with phi^t'_b(ih, il) = Z(n_z(b), ih) R_b(il) of this lane's column b
      const int b = nt * 4 + lc, zr = zb[b];
      const double z0 = Zs[zr * zs + ih];
      double ph[NT];
      ph[0] = z0 * Rb[b * 8];
      if (MODE == 0) {
        const double z1 = Zs[(nzr + zr) * zs + ih];
        ph[1] = z0 * Rb[b * 8 + 1]; ph[2] = z0 * Rb[b * 8 + 2]; ph[3] = z1 * Rb[b * 8];
      }
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int t2 = 0; t2 < NT; t2++) { acc[t][t2][0] += C[t][0] * ph[t2]; acc[t][t2][1] += C[t][1] * ph[t2]; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&t_empty[tg]);
    if (++tg == ntst) { tg = 0; tph ^= 1; }
    if (d.flags & 1) {
      // end of the (s, s') sweep: reduce over the 4 lanes of a grid point (and over the two warps of a shared m-tile,
      // fixed order), write, restart
      const int il = il0 + ilc;
      const bool ok = ih < S.ngh && il < S.ngl;
      constexpr int ndd = NT * NT * 8;
      double* __restrict__ out = (MODE ? g.dd_kap : g.dd_rho) + ((size_t)za * 2 + q) * ndd * g.basis.nghl + (size_t)il * S.ngh + ih;
      double* xb = reinterpret_cast<double*>(smem + L.off_x) + (size_t)((m - first_split) * 8 + lr) * 32;
      double v[NT * NT * 2];
#pragma unroll
      for (int e = 0; e < NT * NT * 2; e++) {
        double x = (&acc[0][0][0])[e];
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        v[e] = x;
        (&acc[0][0][0])[e] = 0.0;
      }
      if (split && half == 1) {
#pragma unroll
        for (int e = 0; e < NT * NT * 2; e++)
          if ((e & 3) == lc) xb[e] = v[e];
      }
      if (split) named_bar_sync(1 + m - first_split, 64);
      if (half == 0) {
#pragma unroll
        for (int e = 0; e < NT * NT * 2; e++)
          if (ok && (e & 3) == lc) out[(size_t)(((e >> 1) * 4 + d.sweep) * 2 + (e & 1)) * g.basis.nghl] = split ? v[e] + xb[e] : v[e];
      }
      if (split) named_bar_sync(1 + m - first_split, 64);
    }
  }
}

void launch_sf_pack(const HamArgs& a, cudaStream_t stream) {
  const SfDev& S = a.sf;
  const int maxsteps = std::max(std::max(S.nsteps[0], S.nsteps[1]), std::max(S.nsteps[2], S.nsteps[3]));
  if (a.nactive > 0 && maxsteps > 0) sf_pack_kernel<<<dim3(maxsteps, 4, a.nactive), 256, 0, stream>>>(a);
}

void launch_density_sf(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  if (a.sf2.enabled) { launch_density_sf2(a, stream); return; }
  const SfDev& S = a.sf;
  const SfDensLayout L = make_dens_layout(S);
  static PerDeviceMax attr_bytes;
  if (attr_bytes.raise(L.total)) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_density_kernel<0, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_density_kernel<1, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_density_kernel<0, 68, 12, 52>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_density_kernel<1, 68, 12, 52>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  }
  const int maxsteps = std::max(std::max(S.nsteps[0], S.nsteps[1]), std::max(S.nsteps[2], S.nsteps[3]));
  if (maxsteps > 0) sf_pack_kernel<<<dim3(maxsteps, 4, a.nactive), 256, 0, stream>>>(a);
  const int nilp = (S.ngl + 1) / 2;
  const dim3 grid(nilp, 2, a.nactive);
  // rho and kappa are independent: the kappa pass runs on a side stream, so its CTAs fill the SMs the rho pass leaves
  // idle (its tail at full batch, most of the GPU once few points are still active)
  SideStreams& ss = *a.side;
  ss.fork_from(stream, 1);
  if (L.ts == 68 && S.kpad_max == 12 && S.zs == 52) {          // 40-point Gauss-Hermite grid, up to 22 shells
    sf_density_kernel<0, 68, 12, 52><<<grid, SF_THREADS, L.total, stream>>>(a, L);
    sf_density_kernel<1, 68, 12, 52><<<grid, SF_THREADS, L.total, ss.s[0]>>>(a, L);
  } else {
    sf_density_kernel<0, 0, 0, 0><<<grid, SF_THREADS, L.total, stream>>>(a, L);
    sf_density_kernel<1, 0, 0, 0><<<grid, SF_THREADS, L.total, ss.s[0]>>>(a, L);
  }
  ss.join_to(stream, 1);
}

// ================================================================================================
// projection
// ================================================================================================
// The unit of work is a PAIR TASK: one DMMA n-tile (4 columns b of one spin segment, re/im interleaved) of one
// (block, a spin segment) output tile, for all il of the CTA's split.  A CTA runs 8 pair tasks that share the spin
// combination (sa, sb) -- and therefore the field tensor of every il -- on 8 independent warp pairs: the phases of a
// pair (G build / DMMA / accumulation) are separated by 64-thread named barriers only, so pairs drift against each
// other and FP64 FMA work of one pair overlaps DMMA work of another; small and large blocks pack on the same SM.
constexpr int SF_PPAIRS = 8;
constexpr int SF_HACC = 12;  // accumulators per lane: rows a = (lane64 >> 3) + 8 i  (na <= 96)

struct SfProjLayout {
  int off_mf, off_pair, off_bar;   // byte offsets: field tensor ring, per-pair regions, mbarriers
  int mf_bytes;                    // bytes of the field tensor of one (il, sa, sb)
  int nst;                         // stages of the field-tensor ring
  int pair_bytes;                  // per pair: G slice | W slice | Ra | ints
  int p_W, p_ra, p_int;            // offsets inside a pair region (G slice at 0)
  int total;
};

template <int MODE>
static SfProjLayout make_proj_layout(const SfDev& S) {
  constexpr int NS = MODE == 0 ? 5 : SF_DIL;
  SfProjLayout L{};
  auto up = [](int x) { return (x + 127) & ~127; };
  int off = up(3 * S.nzrows * S.zs * 8);
  L.mf_bytes = (MODE == 0 ? SF_MFP : SF_DIL) * S.kih * 16;
  const int na_pad = (S.na_max + 7) & ~7;
  const int g_bytes = std::max(NS * S.kih * 8 * 8, 8 * (na_pad + 1) * 8);    // G slice, reused to transpose the output
  L.p_W = up(g_bytes);
  L.p_ra = L.p_W + 2 * 4 * 8 * 8 * 8;
  L.p_int = L.p_ra + up(na_pad * 32);
  L.pair_bytes = up(L.p_int + (2 * na_pad + 8) * 4);
  L.off_pair = off; off += SF_PPAIRS * L.pair_bytes;
  L.off_bar = off; off += 128;
  L.off_mf = off;
  L.nst = std::max(2, std::min(3, (227 * 1024 - off) / up(L.mf_bytes)));
  off += L.nst * up(L.mf_bytes);
  L.total = off;
  return L;
}

// MODE 0: mf -> h;  MODE 1: pf -> Delta
// KIH_, ZS_ (0 = run-time value): padded Gauss-Hermite count and row stride of the z tables as compile-time constants
// for the common grid (see sf_density_kernel)
// HACC_: accumulators per lane (8 rows a each): 6 when no spin segment has more than 48 states, else SF_HACC
template <int MODE, int KIH_, int ZS_, int HACC_>
__global__ void __launch_bounds__(SF_THREADS, 1) sf_projection_kernel(HamArgs g, SfProjLayout L, int q) {
  constexpr int NS = MODE == 0 ? 5 : SF_DIL;   // G slices per iteration: derivative types (h) / Gauss-Laguerre nodes (Delta)
  constexpr int NW = 4;                        // W planes per iteration: radial factor types (h) / nodes (Delta)
  static_assert(SF_DIL == 4, "the W / R_a planes of the Delta pass hold four nodes");
  extern __shared__ __align__(128) unsigned char smem[];
  const SfDev& S = g.sf;
  const int ksp = blockIdx.y, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;   // the host sizes the grid with a stale upper bound of the active slots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int pr = warp >> 1, h = warp & 1, l64 = h * 32 + lane;       // pair, warp of the pair, lane of the pair
  const SfProjTile td = S.tiles[MODE][q][(size_t)blockIdx.x * SF_PPAIRS + pr];
  const SfProjTile t0 = S.tiles[MODE][q][(size_t)blockIdx.x * SF_PPAIRS];   // carries (sa, sb) of the CTA
  const int zs = ZS_ ? ZS_ : S.zs, nzr = S.nzrows, kih = KIH_ ? KIH_ : S.kih;
  const int na = td.na, nslots = td.nslots;
  const bool active = na > 0;
  double* Zs = reinterpret_cast<double*>(smem);                 // [3][nzrows][zs]
  unsigned char* pbase = smem + L.off_pair + (size_t)pr * L.pair_bytes;
  double* G = reinterpret_cast<double*>(pbase);                 // [NS][kih][8], columns swizzled
  double* W = reinterpret_cast<double*>(pbase + L.p_W);         // [2 parts][4][8 slots][8]
  double* Ras = reinterpret_cast<double*>(pbase + L.p_ra);      // [na][4]
  int* slot_a = reinterpret_cast<int*>(pbase + L.p_int);
  const int na_pad = (S.na_max + 7) & ~7;
  int* p2l_a = slot_a + na_pad;
  int* p2l_b = p2l_a + na_pad;                                  // [4]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + L.off_bar);
  unsigned long long* empty = full + 4;
  const int nst = L.nst, dist = nst - 1;
  const int mf_stride = (L.mf_bytes + 127) & ~127;
  // il range of this split
  const int k_per = (S.ngl + S.ksplit - 1) / S.ksplit;
  const int k0 = ksp * k_per, k1 = min(S.ngl, k0 + k_per);
  const int nit = MODE == 0 ? max(0, k1 - k0) : (max(0, k1 - k0) + SF_DIL - 1) / SF_DIL;   // Delta: SF_DIL nodes per iteration

  for (int i = tid; i < 3 * nzr * zs; i += SF_THREADS) Zs[i] = S.zt[i];
  if (active) {
    for (int i = l64; i < na; i += 64) { slot_a[i] = S.slot[td.a_row0 + i]; p2l_a[i] = S.p2l[td.a_row0 + i]; }
    if (l64 < 4) p2l_b[l64] = S.p2l[td.b_row0 + l64];
  }
  if (tid == 0) {
    for (int i = 0; i < nst; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], SF_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // h: mf[il][sa][sb][pair][ih][c];  Delta: pf[sa][sb][il][ih][c] (SF_DIL consecutive nodes are one linear copy; rows
  // beyond the split or the grid are read but masked by zero radial factors)
  const double* __restrict__ mfg =
      MODE == 0 ? g.mf + ((size_t)za * 2 + q) * sf_mf_elems(S.ngl, kih) + (size_t)(t0.sa * 2 + t0.sb) * (L.mf_bytes / 8)
                : g.pf + ((size_t)za * 2 + q) * sf_pf_elems(S.ngl, kih) + (size_t)(t0.sa * 2 + t0.sb) * (S.ngl + SF_DIL) * kih * 2;
  auto issue = [&](int i) {   // field tensor of iteration i -> stage i % nst
    const int sg = i % nst;
    mbar_expect_tx(&full[sg], (unsigned)L.mf_bytes);
    const double* src = MODE == 0 ? mfg + (size_t)(k0 + i) * 4 * (L.mf_bytes / 8) : mfg + (size_t)(k0 + i * SF_DIL) * kih * 2;
    bulk_g2s(smem + L.off_mf + (size_t)sg * mf_stride, src, (unsigned)L.mf_bytes, &full[sg]);
  };
  if (tid == 0)
    for (int i = 0; i < dist && i < nit; i++) issue(i);

  // lane-private constants of the pair task
  const int mt2 = nslots > 8;                                   // two m-tiles of n_z slots: the warps split by m-tile, else by K half
  const int bq = lane & 3;                                      // G phase: column b of this lane
  int zb = 0, zrA = 0;
  if (active) {
    zb = S.zrow[td.b_row0 + bq];
    const int slot = (mt2 ? h * 8 : 0) + lr;
    zrA = slot < nslots ? S.segtab[(size_t)td.seg_a * SF_SEGTAB + 17 + slot] : 0;
  }
  const int noct = (kih + 7) >> 3, oct0 = h == 0 ? 0 : (noct + 1) >> 1, oct1 = h == 0 ? (noct + 1) >> 1 : noct;
  const int nks = kih >> 2;
  const int ks0 = mt2 ? 0 : (h == 0 ? 0 : (nks + 1) >> 1), ks1 = mt2 ? nks : (h == 0 ? (nks + 1) >> 1 : nks);
  const int gsw_st = ((lane >> 3) & 1) << 2, gsw_ld = lr ^ ((lc >> 1) << 2);   // column swizzle of the G slice (store / fragment load)
  // radial factors of iteration 0: R_b of this lane's column (registers), R_a of the rows (3 double2 per lane of the pair)
  double rb[4] = {0.0, 0.0, 0.0, 0.0};
  double2 ra[3];
  auto fetch_r = [&](int i) {   // radial factors of iteration i
    if (!active) return;
    if (MODE == 0) {
      const double* __restrict__ rrow = S.rg + (size_t)(k0 + i) * S.dqp_p * 4;
#pragma unroll
      for (int j = 0; j < 4; j++) rb[j] = rrow[(size_t)(td.b_row0 + bq) * 4 + j];
      const double2* __restrict__ ra2 = reinterpret_cast<const double2*>(rrow + (size_t)td.a_row0 * 4);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int idx = l64 + 64 * r;
        ra[r] = idx < 2 * na ? ra2[idx] : make_double2(0.0, 0.0);
      }
    } else {
      // R0 of SF_DIL nodes: plane j of rb / of the [na][4] row factors is node k0 + i SF_DIL + j (zero beyond the split)
      const int il0 = k0 + i * SF_DIL;
      const size_t ilst = (size_t)S.dqp_p * 4;
#pragma unroll
      for (int j = 0; j < 4; j++) rb[j] = il0 + j < k1 ? S.rg[(size_t)(il0 + j) * ilst + (size_t)(td.b_row0 + bq) * 4] : 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int idx = l64 + 64 * r, a = idx >> 1, j0 = (idx & 1) * 2;
        double x = 0.0, y = 0.0;
        if (idx < 2 * na) {
          const double* __restrict__ ra0 = S.rg + (size_t)(td.a_row0 + a) * 4;
          if (il0 + j0 < k1) x = ra0[(size_t)(il0 + j0) * ilst];
          if (il0 + j0 + 1 < k1) y = ra0[(size_t)(il0 + j0 + 1) * ilst];
        }
        ra[r] = make_double2(x, y);
      }
    }
  };
  if (nit > 0) fetch_r(0);
  double hacc[HACC_];
#pragma unroll
  for (int i = 0; i < HACC_; i++) hacc[i] = 0.0;

  for (int it = 0; it < nit; it++) {
    const int sg = it % nst;
    // refill the ring: the stage of iteration it-1 is free once every warp has left its G phase
    if (tid == 0 && it + dist < nit) {
      if (it >= 1) mbar_wait(&empty[(it + dist) % nst], ((it - 1) / nst) & 1);
      issue(it + dist);
    }
    mbar_wait(&full[sg], (it / nst) & 1);
    if (active) {
      // ---- phase G: G^t(ih, (b,c)) = sum_t' mf^{tt'}(ih) phi^t'_b(ih); lane = (ih of an octet, column b)
      const double2* __restrict__ mfs = reinterpret_cast<const double2*>(smem + L.off_mf + (size_t)sg * mf_stride);
      for (int oc = oct0; oc < oct1; oc++) {
        const int ihg = oc * 8 + (lane >> 2);
        if (ihg < kih) {
          double ph[NS];
          const double z0 = Zs[zb * zs + ihg];
          ph[0] = z0 * rb[0];
          if (MODE == 0) {
            const double z1 = Zs[(nzr + zb) * zs + ihg], z2 = Zs[(2 * nzr + zb) * zs + ihg];
            ph[1] = z0 * rb[1]; ph[2] = z0 * rb[2]; ph[3] = z1 * rb[0]; ph[4] = z2 * rb[0] + z0 * rb[3];
          }
#pragma unroll
          for (int t = 0; t < NS; t++) {
            double gr = 0.0, gi = 0.0;
            if (MODE == 1) {
              const double2 v = mfs[t * kih + ihg];
              const double f = z0 * rb[t];
              gr = v.x * f; gi = v.y * f;
            } else {
#pragma unroll
              for (int t2 = 0; t2 < NS; t2++)
                if (sf_mf_nonzero(t, t2)) {
                  const double2 v = mfs[sf_mf_pair(t, t2) * kih + ihg];
                  gr += v.x * ph[t2]; gi += v.y * ph[t2];
                }
            }
            *reinterpret_cast<double2*>(&G[((size_t)t * kih + ihg) * 8 + ((2 * bq) ^ gsw_st)]) = make_double2(gr, gi);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[sg]);
    if (active) {
      named_bar_sync(1 + pr, 64);
      // radial factors of the rows for this il -> shared memory of the pair (phase C of il-1 is behind the barrier)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int idx = l64 + 64 * r;
        if (idx < 2 * na) reinterpret_cast<double2*>(Ras)[idx] = ra[r];
      }
      // ---- phase W: W^w[slot][(b,c)] = sum_ih Z(slot, ih) G(ih, (b,c)) on the tensor cores
      {
        const double* __restrict__ A0 = Zs + (size_t)zrA * zs + lc;
        const double* __restrict__ gp = G + (size_t)lc * 8 + gsw_ld;
        double C[6][2];
#pragma unroll
        for (int i = 0; i < 12; i++) (&C[0][0])[i] = 0.0;
#pragma unroll 2
        for (int ks = ks0; ks < ks1; ks++) {
          const double a0 = A0[ks * 4];
          const double* __restrict__ gk = gp + (size_t)ks * 32;
          const double g0 = gk[0];
          dmma884(C[0][0], C[0][1], a0, g0);
          if (MODE == 1) {
            const size_t gt = (size_t)kih * 8;
            const double g1 = gk[gt], g2 = gk[2 * gt], g3 = gk[3 * gt];
            dmma884(C[1][0], C[1][1], a0, g1);
            dmma884(C[2][0], C[2][1], a0, g2);
            dmma884(C[3][0], C[3][1], a0, g3);
          } else {
            const double a1 = A0[(size_t)nzr * zs + ks * 4], a2 = A0[(size_t)2 * nzr * zs + ks * 4];
            const size_t gt = (size_t)kih * 8;
            const double g1 = gk[gt], g2 = gk[2 * gt], g3 = gk[3 * gt], g4 = gk[4 * gt];
            dmma884(C[1][0], C[1][1], a0, g1);
            dmma884(C[2][0], C[2][1], a0, g2);
            dmma884(C[3][0], C[3][1], a0, g4);
            dmma884(C[4][0], C[4][1], a1, g3);
            dmma884(C[5][0], C[5][1], a2, g4);
          }
        }
        double* __restrict__ wp = W + ((size_t)(h * 4) * 8 + lr) * 8 + 2 * lc;
        if (MODE == 0) *reinterpret_cast<double2*>(wp) = make_double2(C[0][0] + C[4][0] + C[5][0], C[0][1] + C[4][1] + C[5][1]);
        else *reinterpret_cast<double2*>(wp) = make_double2(C[0][0], C[0][1]);
        *reinterpret_cast<double2*>(wp + 64) = make_double2(C[1][0], C[1][1]);
        *reinterpret_cast<double2*>(wp + 128) = make_double2(C[2][0], C[2][1]);
        *reinterpret_cast<double2*>(wp + 192) = make_double2(C[3][0], C[3][1]);
      }
      // next il's radial factors: issued here, consumed one iteration later
      if (it + 1 < nit) fetch_r(it + 1);
      named_bar_sync(1 + pr, 64);
      // ---- phase C: h[a][(b,c)] += sum_w R^w_a(il) W^w[slot(a)][(b,c)]
      {
        const int col = l64 & 7, ar = l64 >> 3;
#pragma unroll
        for (int i = 0; i < HACC_; i++) {
          const int a = ar + 8 * i;
          if (a < na) {
            const int sl = slot_a[a];
            const double* __restrict__ w0 = W + ((size_t)((mt2 ? sl >> 3 : 0) * 4) * 8 + (sl & 7)) * 8 + col;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
              double x = w0[w * 64];
              if (!mt2) x += w0[(4 + w) * 64];
              s += Ras[a * 4 + w] * x;
            }
            hacc[i] += s;
          }
        }
      }
    }
  }
  // ---- output through the pair's G slice (rows a fastest): partial of this il split, factor 2 applied by the reduction
  if (active) {
    named_bar_sync(1 + pr, 64);
    double* Tr = G;                                     // [col][na | 1]
    const int lda = na | 1;
    const int col = l64 & 7, ar = l64 >> 3;
#pragma unroll
    for (int i = 0; i < HACC_; i++) {
      const int a = ar + 8 * i;
      if (a < na) Tr[col * lda + a] = hacc[i];
    }
    named_bar_sync(1 + pr, 64);
    const size_t pstride = 2 * g.nxy;
    double* __restrict__ part = g.hpart + (((size_t)za * 2 + q) * 2 + MODE) * (size_t)S.ksplit * pstride + (size_t)ksp * pstride;
    for (int idx = l64; idx < 8 * na; idx += 64) {
      const int colx = idx / na, a = idx - colx * na;
      const int lb = p2l_b[colx >> 1];
      if (lb >= 0) part[(size_t)(colx & 1) * g.nxy + td.out_off + p2l_a[a] + (size_t)lb * td.ld] = Tr[colx * lda + a];
    }
  }
}

// sum the split-K partials (fixed order) and scale by 2
__global__ void sf_projection_reduce_kernel(HamArgs g, int ksplit) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y >> 1, is_delta = blockIdx.y & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;   // the host sizes the grid with a stale upper bound of the active slots
  if (e >= 2 * g.nxy) return;
  const int p = g.active[za];
  const double* part = g.hpart + (((size_t)za * 2 + q) * 2 + is_delta) * (size_t)ksplit * 2 * g.nxy;
  double s = 0.0;
  for (int k = 0; k < ksplit; k++) s += part[(size_t)k * 2 * g.nxy + e];
  const int c = e >= g.nxy ? 1 : 0;
  const size_t ee = e - (size_t)c * g.nxy;
  const int quad = is_delta ? g.kap_quad[q] : g.rho_quad[q];
  g.hsp[(((size_t)p * 2 + c) * 4 + quad) * g.nxy + ee] = 2.0 * s;
}

void launch_projection_sf(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  if (a.sf2.enabled) { launch_projection_sf2(a, stream); return; }
  const SfDev& S = a.sf;
  const SfProjLayout L0 = make_proj_layout<0>(S), L1 = make_proj_layout<1>(S);
  static PerDeviceMax attr0, attr1;
  if (attr0.raise(L0.total)) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<0, 0, 0, SF_HACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<0, 40, 52, SF_HACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<0, 40, 52, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0.total));
  }
  if (attr1.raise(L1.total)) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<1, 0, 0, SF_HACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<1, 40, 52, SF_HACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1.total));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf_projection_kernel<1, 40, 52, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1.total));
  }
  const bool common = S.kih == 40 && S.zs == 52;               // 40-point Gauss-Hermite grid
  const bool small = common && S.na_max <= 48;                 // spin segments of at most 48 states (up to 17 shells)
  // four independent launches (h and Delta of both passes): the long ones first, each on its own stream
  SideStreams& ss = *a.side;
  ss.fork_from(stream, 3);
  for (int q = 0; q < 2; q++) {
    const dim3 g0(S.ntiles[0][q] / SF_PPAIRS, S.ksplit, a.nactive);
    cudaStream_t st = q == 0 ? stream : ss.s[0];
    if (S.ntiles[0][q] > 0) {
      if (small) sf_projection_kernel<0, 40, 52, 6><<<g0, SF_THREADS, L0.total, st>>>(a, L0, q);
      else if (common) sf_projection_kernel<0, 40, 52, SF_HACC><<<g0, SF_THREADS, L0.total, st>>>(a, L0, q);
      else sf_projection_kernel<0, 0, 0, SF_HACC><<<g0, SF_THREADS, L0.total, st>>>(a, L0, q);
    }
  }
  for (int q = 0; q < 2; q++) {
    const dim3 g1(S.ntiles[1][q] / SF_PPAIRS, S.ksplit, a.nactive);
    if (S.ntiles[1][q] > 0) {
      if (small) sf_projection_kernel<1, 40, 52, 6><<<g1, SF_THREADS, L1.total, ss.s[1 + q]>>>(a, L1, q);
      else if (common) sf_projection_kernel<1, 40, 52, SF_HACC><<<g1, SF_THREADS, L1.total, ss.s[1 + q]>>>(a, L1, q);
      else sf_projection_kernel<1, 0, 0, SF_HACC><<<g1, SF_THREADS, L1.total, ss.s[1 + q]>>>(a, L1, q);
    }
  }
  ss.join_to(stream, 3);
  dim3 gr((unsigned)((2 * a.nxy + 255) / 256), 4, a.nactive);
  sf_projection_reduce_kernel<<<gr, 256, 0, stream>>>(a, S.ksplit);
}

int sf_density_smem_bytes(const SfDev& S) { return make_dens_layout(S).total; }
int sf_projection_smem_bytes(const SfDev& S) { return std::max(make_proj_layout<0>(S).total, make_proj_layout<1>(S).total); }

}  // namespace pnfam
