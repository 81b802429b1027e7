// Fully factorised density and projection kernels (blocks (b) and (c) of the FAM iteration) for a separable basis.
// Same results as density / meanfield / pairingfield of the reference (exes/pnfam/pnfam_hamiltonian_blas.f90:124-711,
// 717-1169, 1175-1262), as the general-table kernels (hamiltonian.cu) and as the one-sided factorisation
// (hamiltonian_sf.cu); different operation order.
//
// Every table of the harmonic-oscillator basis is a product phi^t_a(ih, il) = Z^m(n_z(a), ih) R^j_a(il)
// (hfbtho_solver.f90:3463-3671) with (m, j) = (0,0) (0,1) (0,2) (1,0) for phi, d/dr, Lambda/r, d/dz and
// (2,0) + (0,3) for the Laplacian.  Using that on BOTH sides of the grid contractions
//
//   density     D^{tt'}_{ss'}(ih,il) = sum_{zr,zr'} Z^m(zr,ih) Z^m'(zr',ih) Pi^{jj'}_{ss'}[zr][zr'][il]
//               Pi^{jj'}_{ss'}[zr][zr'][il] = sum_{a in (s,zr)} sum_{b in (s',zr')} R^j_a(il) rho_ab R^j'_b(il)
//   projection  kt^{jj'}_{sa sb}[zr][zr'][il] = sum_{(t,t') -> (j,j')} sum_ih Z^m(zr,ih) mf^{tt'}_{sa sb}(ih,il) Z^m'(zr',ih)
//               h_ab = 2 sum_il sum_{jj'} R^j_a(il) kt^{jj'}[zr_a][zr_b][il] R^j'_b(il)
//
// (zr = row of the z table, one per distinct n_z of the basis) the work splits into a RADIAL part that touches every
// matrix element once per Gauss-Laguerre node (O(nxy ngl) instead of the reference's O(nxy ngh ngl)) and a part on
// (n_z, n_z') pairs that does not depend on the matrix dimension at all.  At 16 shells the kernels of this file execute
// 1.04 GFLOP per omega point and iteration (ncu, profiles/r02_ncu_kernels.csv) where the GEMM formulation counts 14.6.
// All of it is FP64 FMA work on small operands that live in L1/L2 (on B200 the DFMA and DMMA rates are the same
// 128 flop/clk/SM); the kernels are gather loops without hand-over between warps, bound by load latency and by L1 /
// shared-memory wavefronts, not by the FP64 pipe -- what the measurements on B200 led to:
//   sf2_density_kernel         one CTA per (il, sweep ss', pass, omega): eight lanes own one (zr, zr') entry of Pi and take the
//                              columns that feed it round robin (rho is repacked in that order), the column loop software-
//                              pipelined by hand (one L2 round trip per column and lane otherwise); then the CTA contracts
//                              Pi with the z tables
//   sf2_kappa_density4_kernel  the pairing density (one radial factor): four il per CTA share the element loads
//   sf2_kappa_kernel           one CTA per (il, spin combination, pass, omega): a thread owns the kt entries (zr, 2 zr');
//                              z tables in both index orders (bank conflicts), kt stored [jj][zr'][zr]
//   sf2_radial_kernel          one thread per (two rows a of equal n_z, run of <= 8 columns with equal n_z) of the output
//                              block matrix; component-major radial factors (consecutive lanes = consecutive rows)
// Sums run in a fixed order (no atomics): results are reproducible run to run.
#include <algorithm>
#include <cstdlib>

#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

namespace {
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void cfma(double2& acc, double s, double2 v) { acc.x = fma(s, v.x, acc.x); acc.y = fma(s, v.y, acc.y); }
// index of the (j, j') combination in kt: (0,0) (0,1) (0,2) (0,3) (1,0) (1,1) (1,2) (2,0) (2,1) (2,2) (3,0)
__host__ __device__ constexpr int jj_index(int j, int j2) { return j == 0 ? j2 : (j == 1 ? 4 + j2 : (j == 2 ? 7 + j2 : 10)); }
__host__ __device__ constexpr int mfp(int t, int t2) { return t == 0 ? t2 : (t < 4 ? 5 + (t - 1) * 4 + t2 : 17); }
}  // namespace

// ================================================================================================
// packed copy of rho / kappa in (sweep, zr, zr') order: pk[i] = (re, im) of element el_src[i]
// ================================================================================================
__global__ void __launch_bounds__(256) sf2_pack_kernel(HamArgs g) {
  const SfDev& S = g.sf;
  const Sf2Dev& F = g.sf2;
  const int list = blockIdx.y, kind = list >> 1, q = list & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= F.nelem[list]) return;
  const int p = g.active[za];
  const int quad = kind ? g.kap_quad[q] : g.rho_quad[q];
  const double* __restrict__ src0 = g.rsp + ((size_t)p * 2 + 0) * 4 * g.nxy + (size_t)quad * g.nxy;
  const int e = F.el_src[list][i];
  reinterpret_cast<double2*>(S.pk[kind] + ((size_t)za * 2 + q) * S.pk_stride[kind])[i] = make_double2(src0[e], src0[4 * g.nxy + e]);
}

// ================================================================================================
// density: radial part + contraction with the z tables
// ================================================================================================
// ILS = Gauss-Laguerre nodes per CTA (1 or 2).  With ILS = 2 neighbouring lanes walk the SAME columns for two different
// il: their element loads hit the same sectors (one fetch from L2 and half the L1 wavefronts per il -- every il has to
// stream the whole packed copy, which is what bounds this gather kernel) while the register footprint of a lane is that
// of one il.  The radial factors come from the il-pair table (SfDev::rgp: the two 32-byte records of a row are adjacent).
template <int MODE, int ILS>
__global__ void __launch_bounds__(256 * ILS) sf2_density_kernel(HamArgs g) {
  constexpr int NTHR = 256 * ILS;         // two il: one CTA of 512 threads per SM (its Pi arrays take the room of two CTAs)
  constexpr int NJ = MODE == 0 ? 3 : 1;   // radial factor types of the densities: R0, R1 (d/dr), R2 (Lambda/r)
  constexpr int NT = MODE == 0 ? 4 : 1;   // derivative types: phi, d/dr, Lambda/r, d/dz
  constexpr int KS = 8;                   // lanes sharing one (zr, zr') entry: they stride through its elements
  constexpr int KC = KS / ILS;            // ... of which KC walk different columns
  constexpr int RS = 4 * ILS;             // doubles between two rows of the radial-factor table
  extern __shared__ __align__(16) unsigned char smem[];
  const SfDev& S = g.sf;
  const Sf2Dev& F = g.sf2;
  const int ig = blockIdx.x, sweep = blockIdx.y >> 1, q = blockIdx.y & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;
  const int tid = threadIdx.x;
  const int nzr = F.nzr, npair = nzr * nzr, ngh = S.ngh, list = MODE * 2 + q;
  double2* Pi = reinterpret_cast<double2*>(smem);                        // [ILS][NJ*NJ][npair]
  double* Zs = reinterpret_cast<double*>(Pi + (size_t)ILS * NJ * NJ * npair);  // [2][nzr][ngh]: Z0, Z1
  for (int i = tid; i < 2 * nzr * ngh; i += NTHR) {
    const int m = i / (nzr * ngh), r = i - m * nzr * ngh, zr = r / ngh, ih = r - zr * ngh;
    Zs[i] = S.zt[((size_t)m * nzr + zr) * S.zs + ih];
  }
  for (int i = tid; i < ILS * NJ * NJ * npair; i += NTHR) Pi[i] = make_double2(0.0, 0.0);
  __syncthreads();
  // ---- radial part: Pi^{jj'}[zr][zr'] = sum over the sub-blocks of the group of R^j_a rho_ab R^j'_b at this il, column by
  //      column: t^j = sum_a R^j_a rho_ab (the column is a contiguous run of the packed copy), then t^j R^j'_b.
  //      KS adjacent lanes own one (zr, zr') entry and take its columns round robin (adjacent columns are adjacent in
  //      memory); their partial sums are added in a fixed order (shuffles).  The non-empty entries, heaviest first, are
  //      dealt to the lane groups round by round in snake order: the groups carry nearly equal work (the work of an
  //      entry falls steeply with n_z: the low n_z occur in every block).
  const int* __restrict__ cptr = F.cptr[list] + (size_t)sweep * (npair + 1);
  const int4* __restrict__ cols = F.cols[list];
  const double2* __restrict__ pk = reinterpret_cast<const double2*>(S.pk[MODE] + ((size_t)za * 2 + q) * S.pk_stride[MODE]);
  const int part = tid & (KS - 1), isub = part % ILS, cpart = part / ILS;
  const double* __restrict__ rgl = ILS == 1 ? S.rg + (size_t)ig * S.dqp_p * 4 : S.rgp + (size_t)ig * S.dqp_p * 8 + isub * 4;
  const int* __restrict__ order = F.order[list] + (size_t)sweep * (npair + 1);
  const int nwork = order[0];
  for (int k0 = 0, round = 0; k0 < nwork; k0 += NTHR / KS, round++) {
    const int k = k0 + ((round & 1) ? NTHR / KS - 1 - (tid >> 3) : (tid >> 3));
    const int p = k < nwork ? order[1 + k] : npair;
    double2 acc[NJ][NJ];
#pragma unroll
    for (int i = 0; i < NJ * NJ; i++) (&acc[0][0])[i] = make_double2(0.0, 0.0);
    if (p < npair) {
      // One column per trip.  The loop is bound by the latency of its loads (the packed elements stream from L2, one round
      // trip per column and lane), so it is software-pipelined by hand: the elements of the NEXT column's first row chunk
      // are fetched before the current column is worked on, the column descriptors run two trips ahead.
      const int c1 = cptr[p + 1];
      int c = cptr[p] + cpart;
      int4 cd = make_int4(0, 0, 0, 0), cn = cd;
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = make_double2(0.0, 0.0);
      if (c < c1) {
        cd = __ldg(cols + c);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (u < cd.y) v[u] = __ldg(pk + cd.x + u);
      }
      if (c + KC < c1) cn = __ldg(cols + c + KC);
      while (c < c1) {
        const int cnext = c + KC;
        int4 cnn = make_int4(0, 0, 0, 0);
        if (cnext + KC < c1) cnn = __ldg(cols + cnext + KC);
        double2 vn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) vn[u] = (cnext < c1 && u < cn.y) ? __ldg(pk + cn.x + u) : make_double2(0.0, 0.0);
        const double2* __restrict__ vp = pk + cd.x;           // first element, rows, row of a_0, row of b
        const double* __restrict__ ra = rgl + (size_t)cd.z * RS;
        double2 t[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) t[j] = make_double2(0.0, 0.0);
        {
          double2 r01[4];
          double r2[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool on = u < cd.y;
            r01[u] = on ? ldg2(ra + u * RS) : make_double2(0.0, 0.0);
            r2[u] = (MODE == 0 && on) ? __ldg(ra + u * RS + 2) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            cfma(t[0], r01[u].x, v[u]);
            if (MODE == 0) { cfma(t[1], r01[u].y, v[u]); cfma(t[2], r2[u], v[u]); }
          }
        }
        for (int a0 = 4; a0 < cd.y; a0 += 4) {               // longer columns: the remaining chunks on demand
          double2 w[4], r01[4];
          double r2[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool on = a0 + u < cd.y;
            w[u] = on ? __ldg(vp + a0 + u) : make_double2(0.0, 0.0);
            r01[u] = on ? ldg2(ra + (a0 + u) * RS) : make_double2(0.0, 0.0);
            r2[u] = (MODE == 0 && on) ? __ldg(ra + (a0 + u) * RS + 2) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            cfma(t[0], r01[u].x, w[u]);
            if (MODE == 0) { cfma(t[1], r01[u].y, w[u]); cfma(t[2], r2[u], w[u]); }
          }
        }
        const double2 b01 = ldg2(rgl + (size_t)cd.w * RS);
        const double b2 = MODE == 0 ? __ldg(rgl + (size_t)cd.w * RS + 2) : 0.0;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          cfma(acc[j][0], b01.x, t[j]);
          if (MODE == 0) { cfma(acc[j][1], b01.y, t[j]); cfma(acc[j][2], b2, t[j]); }
        }
        c = cnext; cd = cn; cn = cnn;
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = vn[u];
      }
    }
#pragma unroll
    for (int i = 0; i < NJ * NJ; i++) {                       // sum over the lanes of the same il, fixed order
      double2& v = (&acc[0][0])[i];
      if (ILS == 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, 1); v.y += __shfl_xor_sync(0xffffffffu, v.y, 1); }
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 2); v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 4); v.y += __shfl_xor_sync(0xffffffffu, v.y, 4);
    }
    if (p < npair && cpart == 0) {
#pragma unroll
      for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int j2 = 0; j2 < NJ; j2++) Pi[(size_t)((isub * NJ + j) * NJ + j2) * npair + p] = acc[j][j2];
    }
  }
  __syncthreads();
  // ---- z part: D^{tt'}(ih) = sum_zr Z^{m_t}(zr,ih) sum_zr' Pi^{j_t j_t'}[zr][zr'] Z^{m_t'}(zr',ih); task = (il, ih, t')
  const int2* __restrict__ zrange = F.zrange[list] + (size_t)sweep * nzr;
  constexpr int ndd = NT * NT * 8;
  for (int task = tid; task < ILS * ngh * NT; task += NTHR) {
    const int is = task / (ngh * NT), tr = task - is * ngh * NT;
    const int il = ig * ILS + is;
    if (il >= S.ngl) continue;
    const int t2 = tr / ngh, ih = tr - t2 * ngh;
    const int j2 = (MODE == 0 && t2 < 3) ? t2 : 0, m2 = (MODE == 0 && t2 == 3) ? 1 : 0;
    const double2* __restrict__ Pil = Pi + (size_t)is * NJ * NJ * npair;
    double* __restrict__ out0 = (MODE ? g.dd_kap : g.dd_rho) + ((size_t)za * 2 + q) * ndd * g.basis.nghl + (size_t)il * ngh;
    double2 d[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) d[t] = make_double2(0.0, 0.0);
    for (int zr = 0; zr < nzr; zr++) {
      const int2 rng = zrange[zr];
      if (rng.x >= rng.y) continue;
      double2 x[NJ];
#pragma unroll
      for (int j = 0; j < NJ; j++) x[j] = make_double2(0.0, 0.0);
      const double* __restrict__ zcol = Zs + (size_t)m2 * nzr * ngh + ih;
      const double2* __restrict__ prow = Pil + (size_t)j2 * npair + (size_t)zr * nzr;
      for (int z2 = rng.x; z2 < rng.y; z2++) {
        const double zz = zcol[(size_t)z2 * ngh];
#pragma unroll
        for (int j = 0; j < NJ; j++) cfma(x[j], zz, prow[(size_t)j * NJ * npair + z2]);
      }
      const double z0 = Zs[(size_t)zr * ngh + ih];
      cfma(d[0], z0, x[0]);
      if (MODE == 0) {
        const double z1 = Zs[(size_t)(nzr + zr) * ngh + ih];
        cfma(d[1], z0, x[1]); cfma(d[2], z0, x[2]); cfma(d[3], z1, x[0]);
      }
    }
#pragma unroll
    for (int t = 0; t < NT; t++) {
      double* __restrict__ o = out0 + (size_t)(((t * NT + t2) * 4 + sweep) * 2) * g.basis.nghl + ih;
      o[0] = d[t].x;
      o[g.basis.nghl] = d[t].y;
    }
  }
}

// ================================================================================================
// pairing density, four Gauss-Laguerre nodes per CTA.  kappa -> K needs the plain wave function only (one radial factor
// R0 per row), so one pass over the packed kappa elements can serve several il at once: the element loads -- the traffic
// that bounds these gather kernels (every il re-reads the whole packed copy from L2) -- are shared by four nodes and the
// four R0 of a row come as one 32-byte record (SfDev::r0q).
// ================================================================================================
__global__ void __launch_bounds__(256) sf2_kappa_density4_kernel(HamArgs g) {
  constexpr int KS = 8, NI = 4;
  extern __shared__ __align__(16) unsigned char smem[];
  const SfDev& S = g.sf;
  const Sf2Dev& F = g.sf2;
  const int iq = blockIdx.x, sweep = blockIdx.y >> 1, q = blockIdx.y & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;
  const int tid = threadIdx.x;
  const int nzr = F.nzr, npair = nzr * nzr, ngh = S.ngh, list = 2 + q;
  double2* Pi = reinterpret_cast<double2*>(smem);                        // [NI][npair]
  double* Zs = reinterpret_cast<double*>(Pi + (size_t)NI * npair);      // [nzr][ngh]: Z0
  for (int i = tid; i < nzr * ngh; i += 256) {
    const int zr = i / ngh, ih = i - zr * ngh;
    Zs[i] = S.zt[(size_t)zr * S.zs + ih];
  }
  for (int i = tid; i < NI * npair; i += 256) Pi[i] = make_double2(0.0, 0.0);
  __syncthreads();
  const int* __restrict__ cptr = F.cptr[list] + (size_t)sweep * (npair + 1);
  const int4* __restrict__ cols = F.cols[list];
  const double2* __restrict__ pk = reinterpret_cast<const double2*>(S.pk[1] + ((size_t)za * 2 + q) * S.pk_stride[1]);
  const double* __restrict__ rq = S.r0q + (size_t)iq * S.dqp_p * 4;
  const int part = tid & (KS - 1);
  const int* __restrict__ order = F.order[list] + (size_t)sweep * (npair + 1);
  const int nwork = order[0];
  for (int k0 = 0, round = 0; k0 < nwork; k0 += 256 / KS, round++) {
    const int k = k0 + ((round & 1) ? 256 / KS - 1 - (tid >> 3) : (tid >> 3));
    const int p = k < nwork ? order[1 + k] : npair;
    double2 acc[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) acc[i] = make_double2(0.0, 0.0);
    if (p < npair) {
      // software-pipelined like the rho density: the next column's elements are fetched one trip ahead
      const int c1 = cptr[p + 1];
      int c = cptr[p] + part;
      int4 cd = make_int4(0, 0, 0, 0), cn = cd;
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = make_double2(0.0, 0.0);
      if (c < c1) {
        cd = __ldg(cols + c);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (u < cd.y) v[u] = __ldg(pk + cd.x + u);
      }
      if (c + KS < c1) cn = __ldg(cols + c + KS);
      while (c < c1) {
        const int cnext = c + KS;
        int4 cnn = make_int4(0, 0, 0, 0);
        if (cnext + KS < c1) cnn = __ldg(cols + cnext + KS);
        double2 vn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) vn[u] = (cnext < c1 && u < cn.y) ? __ldg(pk + cn.x + u) : make_double2(0.0, 0.0);
        const double2* __restrict__ vp = pk + cd.x;
        const double* __restrict__ ra = rq + (size_t)cd.z * 4;
        double2 t[NI];
#pragma unroll
        for (int i = 0; i < NI; i++) t[i] = make_double2(0.0, 0.0);
        {
          double2 r01[4], r23[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool on = u < cd.y;
            r01[u] = on ? ldg2(ra + u * 4) : make_double2(0.0, 0.0);
            r23[u] = on ? ldg2(ra + u * 4 + 2) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) { cfma(t[0], r01[u].x, v[u]); cfma(t[1], r01[u].y, v[u]); cfma(t[2], r23[u].x, v[u]); cfma(t[3], r23[u].y, v[u]); }
        }
        for (int a0 = 4; a0 < cd.y; a0 += 4) {
          double2 w[4], r01[4], r23[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool on = a0 + u < cd.y;
            w[u] = on ? __ldg(vp + a0 + u) : make_double2(0.0, 0.0);
            r01[u] = on ? ldg2(ra + (a0 + u) * 4) : make_double2(0.0, 0.0);
            r23[u] = on ? ldg2(ra + (a0 + u) * 4 + 2) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) { cfma(t[0], r01[u].x, w[u]); cfma(t[1], r01[u].y, w[u]); cfma(t[2], r23[u].x, w[u]); cfma(t[3], r23[u].y, w[u]); }
        }
        const double2 b01 = ldg2(rq + (size_t)cd.w * 4), b23 = ldg2(rq + (size_t)cd.w * 4 + 2);
        cfma(acc[0], b01.x, t[0]); cfma(acc[1], b01.y, t[1]); cfma(acc[2], b23.x, t[2]); cfma(acc[3], b23.y, t[3]);
        c = cnext; cd = cn; cn = cnn;
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = vn[u];
      }
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
      double2& v = acc[i];
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 1); v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 2); v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 4); v.y += __shfl_xor_sync(0xffffffffu, v.y, 4);
    }
    if (p < npair && part == 0) {
#pragma unroll
      for (int i = 0; i < NI; i++) Pi[(size_t)i * npair + p] = acc[i];
    }
  }
  __syncthreads();
  // ---- z part: K(ih, il) = sum_zr Z0(zr,ih) sum_zr' Pi[il][zr][zr'] Z0(zr',ih); task = (il of the group, ih)
  const int2* __restrict__ zrange = F.zrange[list] + (size_t)sweep * nzr;
  for (int task = tid; task < ngh * NI; task += 256) {
    const int i = task / ngh, ih = task - i * ngh, il = iq * NI + i;
    if (il >= S.ngl) continue;
    double2 d = make_double2(0.0, 0.0);
    for (int zr = 0; zr < nzr; zr++) {
      const int2 rng = zrange[zr];
      if (rng.x >= rng.y) continue;
      double2 x = make_double2(0.0, 0.0);
      const double2* __restrict__ prow = Pi + (size_t)i * npair + (size_t)zr * nzr;
      for (int z2 = rng.x; z2 < rng.y; z2++) cfma(x, Zs[(size_t)z2 * ngh + ih], prow[z2]);
      cfma(d, Zs[(size_t)zr * ngh + ih], x);
    }
    double* __restrict__ o = g.dd_kap + ((size_t)za * 2 + q) * 8 * g.basis.nghl + (size_t)il * ngh + (size_t)(sweep * 2) * g.basis.nghl + ih;
    o[0] = d.x;
    o[g.basis.nghl] = d.y;
  }
}

void launch_density_sf2(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  const SfDev& S = a.sf;
  const int nzr = a.sf2.nzr, npair = nzr * nzr;
  // two il per CTA (512 threads, one CTA per SM) when its two sets of Pi arrays fit: with the column loop software-
  // pipelined the shared element fetches pay (1.93 against 2.18 ms per launch at 16 shells, 64 points; before the
  // pipelining the variant was slower).  PNFAM_B200_DENSITY_ILS=1 keeps one il per CTA.
  static const int ils_knob = getenv("PNFAM_B200_DENSITY_ILS") ? atoi(getenv("PNFAM_B200_DENSITY_ILS")) : 2;
  const size_t zbytes = (size_t)2 * nzr * S.ngh * 8;
  const int ils = (ils_knob >= 2 && (size_t)18 * npair * 16 + zbytes <= 220 * 1024) ? 2 : 1;
  const size_t sm0 = (size_t)ils * 9 * npair * 16 + zbytes, sm4 = (size_t)4 * npair * 16 + (size_t)nzr * S.ngh * 8;
  static PerDeviceMax attr;
  if (attr.raise(std::max(sm0, sm4))) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf2_density_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(sm0, sm4)));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf2_density_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(sm0, sm4)));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf2_kappa_density4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(sm0, sm4)));
  }
  const int nel = std::max(std::max(a.sf2.nelem[0], a.sf2.nelem[1]), std::max(a.sf2.nelem[2], a.sf2.nelem[3]));
  if (nel > 0) sf2_pack_kernel<<<dim3((nel + 255) / 256, 4, a.nactive), 256, 0, stream>>>(a);
  SideStreams& ss = *a.side;
  ss.fork_from(stream, 1);
  if (ils == 2) sf2_density_kernel<0, 2><<<dim3((S.ngl + 1) / 2, 8, a.nactive), 512, sm0, stream>>>(a);
  else sf2_density_kernel<0, 1><<<dim3(S.ngl, 8, a.nactive), 256, sm0, stream>>>(a);
  sf2_kappa_density4_kernel<<<dim3((S.ngl + 3) / 4, 8, a.nactive), 256, sm4, ss.s[0]>>>(a);
  ss.join_to(stream, 1);
}

// ================================================================================================
// projection, z part: kt^{jj'}[zr][zr'] of one (il, sa, sb)
// ================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256, 2) sf2_kappa_kernel(HamArgs g) {
  constexpr int NP = MODE == 0 ? SF_MFP : 1;       // field-tensor pairs
  constexpr int NJJ = MODE == 0 ? SF2_NJJ : 1;
  extern __shared__ __align__(16) unsigned char smem[];
  const SfDev& S = g.sf;
  const Sf2Dev& F = g.sf2;
  const int il = blockIdx.x, s4 = blockIdx.y >> 1, q = blockIdx.y & 1, za = blockIdx.z;
  if (g.ctrl && za >= g.ctrl->nactive) return;
  const int tid = threadIdx.x;
  const int nzr = F.nzr, npair = nzr * nzr, ngh = S.ngh, kih = S.kih;
  double2* mfs = reinterpret_cast<double2*>(smem);                     // [NP][ngh]
  double* Zs = reinterpret_cast<double*>(mfs + (size_t)NP * ngh);      // [3][nzr][ngh]
  const double* __restrict__ src =
      MODE == 0 ? g.mf + ((size_t)za * 2 + q) * sf_mf_elems(S.ngl, kih) + ((size_t)il * 4 + s4) * SF_MFP * kih * 2
                : g.pf + ((size_t)za * 2 + q) * sf_pf_elems(S.ngl, kih) + ((size_t)s4 * (S.ngl + SF_DIL) + il) * kih * 2;
  for (int i = tid; i < NP * ngh; i += 256) {
    const int pr = i / ngh, ih = i - pr * ngh;
    mfs[i] = *reinterpret_cast<const double2*>(src + ((size_t)pr * kih + ih) * 2);
  }
  constexpr int NM = MODE == 0 ? 3 : 1;
  // two copies of the z tables: [m][zr][ih] for the row index (nearly uniform over a warp) and [m][ih][zr'] for the column
  // index (consecutive lanes = consecutive zr': conflict-free; the [zr][ih] order would put 32 lanes on two banks)
  const int nzp = nzr | 1;
  double* Zt = Zs + (size_t)NM * nzr * ngh;                            // [NM][ngh][nzp]
  for (int i = tid; i < NM * nzr * ngh; i += 256) {
    const int m = i / (nzr * ngh), r = i - m * nzr * ngh, zr = r / ngh, ih = r - zr * ngh;
    const double v = S.zt[((size_t)m * nzr + zr) * S.zs + ih];
    Zs[i] = v;
    Zt[((size_t)m * ngh + ih) * nzp + zr] = v;
  }
  __syncthreads();
  const unsigned char* __restrict__ need = F.need[MODE][q] + (size_t)s4 * npair;
  double* __restrict__ out = F.kt[MODE] + ((size_t)za * 2 + q) * sf2_kt_elems(MODE, S.ngl, nzr) + ((size_t)s4 * S.ngl + il) * npair * NJJ * 2;
  // kt is stored [jj][zr'][zr] (zr fastest): the radial kernel's lanes are consecutive rows a -- a few neighbouring zr at
  // one zr' -- so its eleven loads per il touch two or three lines instead of one line per distinct zr.  Consecutive
  // lanes here therefore differ in zr (the [ih][zr] copy of the z tables), zr' is nearly uniform over a warp.
  // A thread works on TWO columns zr' (2k, 2k+1) of its row zr: the kernel is bound by shared-memory wavefronts, and the
  // field-tensor broadcasts (18 of the 24 loads per ih) then serve two kt entries.
  const int nz2 = (nzr + 1) >> 1;
  for (int pp = tid; pp < nz2 * nzr; pp += 256) {
    const int zk = pp / nzr, zr = pp - zk * nzr, z2a = 2 * zk, z2b = min(2 * zk + 1, nzr - 1);
    const bool ona = need[zr * nzr + z2a] != 0, onb = 2 * zk + 1 < nzr && need[zr * nzr + z2b] != 0;
    if (!ona && !onb) continue;
    double2 acc[NJJ], acd[NJJ];
#pragma unroll
    for (int i = 0; i < NJJ; i++) acc[i] = acd[i] = make_double2(0.0, 0.0);
    const double* __restrict__ za_ = Zt + zr;
    const double* __restrict__ zb_ = Zs + (size_t)z2a * ngh;
    const double* __restrict__ zc_ = Zs + (size_t)z2b * ngh;
    const size_t ms = (size_t)nzr * ngh, mt = (size_t)ngh * nzp;
    for (int ih = 0; ih < ngh; ih++) {
      const double a0 = za_[(size_t)ih * nzp], b0 = zb_[ih], c0 = zc_[ih];
      const double p00 = a0 * b0, q00 = a0 * c0;
      if (MODE == 1) {
        const double2 m = mfs[ih];
        cfma(acc[0], p00, m); cfma(acd[0], q00, m);
      } else {
        const double a1 = za_[mt + (size_t)ih * nzp], a2 = za_[2 * mt + (size_t)ih * nzp];
        const double b1 = zb_[ms + ih], b2 = zb_[2 * ms + ih], c1 = zc_[ms + ih], c2 = zc_[2 * ms + ih];
        const double p01 = a0 * b1, p10 = a1 * b0, p11 = a1 * b1, p02 = a0 * b2, p20 = a2 * b0;
        const double q01 = a0 * c1, q10 = a1 * c0, q11 = a1 * c1, q02 = a0 * c2, q20 = a2 * c0;
        // (t, t') with t, t' in {phi, d/dr, Lambda/r}: Z0 Z0', radial pair (t, t')
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
          for (int t2 = 0; t2 < 3; t2++) { const double2 m = mfs[mfp(t, t2) * ngh + ih]; cfma(acc[jj_index(t, t2)], p00, m); cfma(acd[jj_index(t, t2)], q00, m); }
        // d/dz on one side: Z1, radial factor R0
#pragma unroll
        for (int t = 0; t < 3; t++) {
          { const double2 m = mfs[mfp(t, 3) * ngh + ih]; cfma(acc[jj_index(t, 0)], p01, m); cfma(acd[jj_index(t, 0)], q01, m); }
          { const double2 m = mfs[mfp(3, t) * ngh + ih]; cfma(acc[jj_index(0, t)], p10, m); cfma(acd[jj_index(0, t)], q10, m); }
        }
        { const double2 m = mfs[mfp(3, 3) * ngh + ih]; cfma(acc[jj_index(0, 0)], p11, m); cfma(acd[jj_index(0, 0)], q11, m); }
        // Laplacian = Z2 R0 + Z0 R3, only next to the plain wave function
        { const double2 m = mfs[mfp(0, 4) * ngh + ih];
          cfma(acc[jj_index(0, 0)], p02, m); cfma(acc[jj_index(0, 3)], p00, m); cfma(acd[jj_index(0, 0)], q02, m); cfma(acd[jj_index(0, 3)], q00, m); }
        { const double2 m = mfs[mfp(4, 0) * ngh + ih];
          cfma(acc[jj_index(0, 0)], p20, m); cfma(acc[jj_index(3, 0)], p00, m); cfma(acd[jj_index(0, 0)], q20, m); cfma(acd[jj_index(3, 0)], q00, m); }
      }
    }
    double2* __restrict__ o = reinterpret_cast<double2*>(out);
    if (ona) {
#pragma unroll
      for (int i = 0; i < NJJ; i++) o[(size_t)i * npair + z2a * nzr + zr] = acc[i];
    }
    if (onb) {
#pragma unroll
      for (int i = 0; i < NJJ; i++) o[(size_t)i * npair + z2b * nzr + zr] = acd[i];
    }
  }
}

// ================================================================================================
// projection, radial part: one thread = one row a x a run of <= SF2_RUN columns with equal n_z
// ================================================================================================
#ifndef SF2_RADIAL_UNROLL
#define SF2_RADIAL_UNROLL 2
#endif
#ifndef SF2_RADIAL_CTAS
#define SF2_RADIAL_CTAS 3
#endif
template <int MODE>
__global__ void __launch_bounds__(128, SF2_RADIAL_CTAS) sf2_radial_kernel(HamArgs g, int q) {
  constexpr int NJJ = MODE == 0 ? SF2_NJJ : 1;
  const SfDev& S = g.sf;
  const Sf2Dev& F = g.sf2;
  const int za = blockIdx.y;
  if (g.ctrl && za >= g.ctrl->nactive) return;
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= F.ntasks[MODE][q]) return;
  const Sf2Task t = F.tasks[MODE][q][k];
  const int nzr = F.nzr, npair = nzr * nzr, ngl = S.ngl;
  const int zra = S.zrow[t.pa], zrb = S.zrow[t.pb0];
  const size_t kstride = (size_t)npair * NJJ * 2, rstride = (size_t)S.dqp_p * 4;
  const double* __restrict__ kp = F.kt[MODE] + ((size_t)za * 2 + q) * sf2_kt_elems(MODE, ngl, nzr) + (size_t)t.sasb * ngl * kstride +
                                  ((size_t)zrb * nzr + zra) * 2;     // [jj][zr'][zr]
  const size_t kj = (size_t)npair * 2;                                // doubles between two jj
  const double* __restrict__ ra = S.rg + (size_t)t.pa * 4;
  const double* __restrict__ rb = S.rg + (size_t)t.pb0 * 4;
  // two rows of equal n_z per thread: the kt entries and the column factors of an il are fetched once for both (the
  // kernel is bound by L1 throughput); a single-row task runs the second row with zero factors
  const double f2 = t.na > 1 ? 1.0 : 0.0;
  const size_t np = (size_t)S.dqp_p;
  const double* __restrict__ rt = S.rgt + t.pa;             // component-major radial factors: [il][4][row]
  const int o2 = t.na > 1 ? 1 : 0;
  double2 acc[SF2_RUN], acd[SF2_RUN];
#pragma unroll
  for (int c = 0; c < SF2_RUN; c++) acc[c] = acd[c] = make_double2(0.0, 0.0);
  for (int il = 0; il < ngl; il++, kp += kstride, rt += 4 * np, rb += rstride) {
    const double2 a01 = make_double2(__ldg(rt), __ldg(rt + np));
    double2 c01 = make_double2(__ldg(rt + o2), __ldg(rt + np + o2));
    c01.x *= f2; c01.y *= f2;
    if (MODE == 1) {
      const double2 kk = ldg2(kp);
      const double2 v = make_double2(a01.x * kk.x, a01.x * kk.y), w = make_double2(c01.x * kk.x, c01.x * kk.y);
#pragma unroll
      for (int c = 0; c < SF2_RUN; c++)
        if (c < t.nb) { const double b0 = __ldg(rb + c * 4); cfma(acc[c], b0, v); cfma(acd[c], b0, w); }
    } else {
      const double2 a23 = make_double2(__ldg(rt + 2 * np), __ldg(rt + 3 * np));
      double2 c23 = make_double2(__ldg(rt + 2 * np + o2), __ldg(rt + 3 * np + o2));
      c23.x *= f2; c23.y *= f2;
      double2 v0 = make_double2(0.0, 0.0), v1 = v0, v2 = v0, v3 = v0, w0 = v0, w1 = v0, w2 = v0, w3 = v0;
      // V^{j'} = sum_j R^j_a kt^{jj'} for both rows
      { const double2 k = ldg2(kp + kj * jj_index(0, 0)); cfma(v0, a01.x, k); cfma(w0, c01.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(1, 0)); cfma(v0, a01.y, k); cfma(w0, c01.y, k); }
      { const double2 k = ldg2(kp + kj * jj_index(2, 0)); cfma(v0, a23.x, k); cfma(w0, c23.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(3, 0)); cfma(v0, a23.y, k); cfma(w0, c23.y, k); }
      { const double2 k = ldg2(kp + kj * jj_index(0, 1)); cfma(v1, a01.x, k); cfma(w1, c01.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(1, 1)); cfma(v1, a01.y, k); cfma(w1, c01.y, k); }
      { const double2 k = ldg2(kp + kj * jj_index(2, 1)); cfma(v1, a23.x, k); cfma(w1, c23.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(0, 2)); cfma(v2, a01.x, k); cfma(w2, c01.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(1, 2)); cfma(v2, a01.y, k); cfma(w2, c01.y, k); }
      { const double2 k = ldg2(kp + kj * jj_index(2, 2)); cfma(v2, a23.x, k); cfma(w2, c23.x, k); }
      { const double2 k = ldg2(kp + kj * jj_index(0, 3)); cfma(v3, a01.x, k); cfma(w3, c01.x, k); }
#pragma unroll
      for (int c = 0; c < SF2_RUN; c++)
        if (c < t.nb) {
          const double2 b01 = ldg2(rb + c * 4), b23 = ldg2(rb + c * 4 + 2);
          cfma(acc[c], b01.x, v0); cfma(acc[c], b01.y, v1); cfma(acc[c], b23.x, v2); cfma(acc[c], b23.y, v3);
          cfma(acd[c], b01.x, w0); cfma(acd[c], b01.y, w1); cfma(acd[c], b23.x, w2); cfma(acd[c], b23.y, w3);
        }
    }
  }
  const int p = g.active[za];
  const int quad = MODE ? g.kap_quad[q] : g.rho_quad[q];
  double* __restrict__ ore = g.hsp + (((size_t)p * 2 + 0) * 4 + quad) * g.nxy + t.out_base + S.p2l[t.pa];
  double* __restrict__ oim = ore + 4 * g.nxy;
  const int d2 = t.na > 1 ? S.p2l[t.pa + 1] - S.p2l[t.pa] : 0;
#pragma unroll
  for (int c = 0; c < SF2_RUN; c++)
    if (c < t.nb) {
      const size_t e = (size_t)S.p2l[t.pb0 + c] * t.ld;
      ore[e] = 2.0 * acc[c].x;
      oim[e] = 2.0 * acc[c].y;
      if (t.na > 1) { ore[e + d2] = 2.0 * acd[c].x; oim[e + d2] = 2.0 * acd[c].y; }
    }
}

void launch_projection_sf2(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  const SfDev& S = a.sf;
  const Sf2Dev& F = a.sf2;
  const int nzr = F.nzr;
  const size_t sm0 = (size_t)SF_MFP * S.ngh * 16 + (size_t)3 * (nzr + (nzr | 1)) * S.ngh * 8, sm1 = (size_t)S.ngh * 16 + (size_t)(nzr + (nzr | 1)) * S.ngh * 8;
  static PerDeviceMax attr;
  if (sm0 > 48 * 1024 && attr.raise(sm0))
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(sf2_kappa_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0));
  // the mean field on the caller's stream, the pairing field (a tenth of the work) on a side stream
  SideStreams& ss = *a.side;
  ss.fork_from(stream, 1);
  const dim3 gk(S.ngl, 8, a.nactive);
  sf2_kappa_kernel<0><<<gk, 256, sm0, stream>>>(a);
  sf2_kappa_kernel<1><<<gk, 256, sm1, ss.s[0]>>>(a);
  for (int q = 0; q < 2; q++) {
    if (F.ntasks[0][q] > 0) sf2_radial_kernel<0><<<dim3((F.ntasks[0][q] + 127) / 128, a.nactive), 128, 0, stream>>>(a, q);
    if (F.ntasks[1][q] > 0) sf2_radial_kernel<1><<<dim3((F.ntasks[1][q] + 127) / 128, a.nactive), 128, 0, ss.s[0]>>>(a, q);
  }
  ss.join_to(stream, 1);
}

int sf2_smem_bytes(const SfDev& S) { return 9 * S.nzrows * S.nzrows * 16 + 2 * S.nzrows * S.ngh * 8; }   // one il per CTA

}  // namespace pnfam
