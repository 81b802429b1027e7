// Blocks (b) and (c) of the FAM iteration: perturbed densities on the Gauss-Hermite x Gauss-Laguerre
// grid, pointwise Skyrme / pairing fields, and the grid -> HO projection of the induced fields.
// Replaces calc_hamiltonian = density -> meanfield -> pairingfield
// (exes/pnfam/pnfam_hamiltonian_blas.f90:51-74, 124-711, 717-1169, 1175-1262).
//
// B200-first formulation (not the reference's): the reference materialises
//   wfa_rhoab^t_s(r,b) = sum_{a in s} phi^t_a(r) rho_ab         (16 thin DGEMMs per block, 4 Ng x N scratch arrays)
// and then streams ~20 Ng x N arrays through serial pointwise loops.  Here the grid point is the
// outer tile: one CTA owns RT=16 grid points, forms the same product tile-by-tile on the FP64
// tensor cores (DMMA) and contracts it IMMEDIATELY with phi^t'_b(r) in the epilogue, so that only the
// 64 bilinear forms
//   D^{t t'}_{s s'}(r) = sum_{a in s, b in s'} phi^t_a(r) rho_ab phi^t'_b(r)
// ever leave the SM (128 doubles per grid point instead of 16*N).  All 24 local densities are fixed
// linear combinations of D; the field tensor mf(t,t',s,s')(r) is pointwise; the projection
//   h_ab = 2 sum_r sum_{t t'} phi^t_a(r) mf^{t t'}_{s_a s_b}(r) phi^t'_b(r)
// builds G^t(r,b) = sum_t' mf^{t t'} phi^t'_b on the fly in shared memory and contracts over (r,t)
// with DMMA, so the 20 Ng x N "hpsi" arrays of the reference never exist either.
#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

constexpr int AC = 64;    // basis states (contraction index) per shared-memory chunk
constexpr int BC = 32;    // columns b per chunk (8 DMMA n-tiles of 4 b x {re,im})
constexpr int RHS = 68;   // padded row stride of the interleaved (b,c) chunk of rho

// ================================================================================================
// density: D^{t t'}_{s s'}(r)
// ================================================================================================
template <int NT>
struct DensSmem {
  double a[NT][AC][RS];
  double b[NT][BC][RS];
  double rho[AC][RHS];
};

template <int NT>
__global__ void __launch_bounds__(256) density_kernel(HamArgs g, int is_kappa) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DensSmem<NT>& sm = *reinterpret_cast<DensSmem<NT>*>(smem_raw);
  constexpr int CG = (NT == 4) ? 1 : 4;       // column groups (warps sharing an r-half split the n-tiles)
  constexpr int NTW = 8 / CG;                 // n-tiles per warp per chunk
  const int tile = blockIdx.x, q = blockIdx.y, za = blockIdx.z;
  const int p = g.active[za];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  const int tw = (NT == 4) ? (warp & 3) : 0;
  const int rh = (NT == 4) ? (warp >> 2) : (warp & 1);
  const int cg = (NT == 4) ? 0 : (warp >> 1);
  const DevBasis& B = g.basis;
  const DevBlockStruct st = is_kappa ? g.kap_in[q] : g.rho_in[q];
  const int quad = is_kappa ? g.kap_quad[q] : g.rho_quad[q];
  const double* __restrict__ rre = g.rsp + ((size_t)p * 2 + 0) * 4 * g.nxy + (size_t)quad * g.nxy;
  const double* __restrict__ rim = g.rsp + ((size_t)p * 2 + 1) * 4 * g.nxy + (size_t)quad * g.nxy;
  const double* __restrict__ phit = B.phi + (size_t)tile * NTYPE * B.dqp * RT;
  // type index in the global table for local type index: rho uses wf, dr, dp, dz = 0..3; kappa uses wf only
  double acc[2][2][NT][2];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int sp = 0; sp < 2; sp++)
#pragma unroll
      for (int t = 0; t < NT; t++) acc[s][sp][t][0] = acc[s][sp][t][1] = 0.0;

  for (int ix = 0; ix < B.nb; ix++) {
    const int iy = st.r2c[ix];
    if (iy < 0) continue;
    const int di = B.db[ix], dj = B.db[iy], nui = B.nsu[ix], nuj = B.nsu[iy];
    const int ia = B.isstart[ix], ib = B.isstart[iy];
    const size_t off = st.r2m[ix];
#pragma unroll
    for (int sp = 0; sp < 2; sp++) {
      const int b_lo = sp == 0 ? 0 : nuj, b_hi = sp == 0 ? nuj : dj;
      for (int bc0 = b_lo; bc0 < b_hi; bc0 += BC) {
        const int nbc = min(BC, b_hi - bc0), nbc4 = (nbc + 3) & ~3;
        __syncthreads();
        for (int idx = threadIdx.x; idx < NT * nbc4 * RT; idx += 256) {
          const int rr = idx & (RT - 1), bl = (idx / RT) % nbc4, t = idx / (RT * nbc4);
          sm.b[t][bl][rr] = bl < nbc ? phit[((size_t)t * B.dqp + ib + bc0 + bl) * RT + rr] : 0.0;
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const int a_lo = s == 0 ? 0 : nui, a_hi = s == 0 ? nui : di;
          if (a_hi <= a_lo) continue;
          double C[NTW][2];
#pragma unroll
          for (int j = 0; j < NTW; j++) C[j][0] = C[j][1] = 0.0;
          for (int ac0 = a_lo; ac0 < a_hi; ac0 += AC) {
            const int nac = min(AC, a_hi - ac0), nac4 = (nac + 3) & ~3;
            __syncthreads();
            for (int idx = threadIdx.x; idx < NT * nac4 * RT; idx += 256) {
              const int rr = idx & (RT - 1), al = (idx / RT) % nac4, t = idx / (RT * nac4);
              sm.a[t][al][rr] = al < nac ? phit[((size_t)t * B.dqp + ia + ac0 + al) * RT + rr] : 0.0;
            }
            for (int idx = threadIdx.x; idx < nac4 * nbc4; idx += 256) {
              const int al = idx % nac4, bl = idx / nac4;
              double vr = 0.0, vi = 0.0;
              if (al < nac && bl < nbc) {
                const size_t e = off + (size_t)(ac0 + al) + (size_t)(bc0 + bl) * di;
                vr = rre[e]; vi = rim[e];
              }
              sm.rho[al][2 * bl] = vr;
              sm.rho[al][2 * bl + 1] = vi;
            }
            __syncthreads();
            const int ksteps = nac4 >> 2;
            for (int ks = 0; ks < ksteps; ks++) {
              const double af = sm.a[tw][ks * 4 + lc][rh * 8 + lr];
#pragma unroll
              for (int j = 0; j < NTW; j++) {
                const int nt = cg + CG * j;
                if (nt * 4 < nbc4) {
                  const double bf = sm.rho[ks * 4 + lc][nt * 8 + lr];
                  dmma884(C[j][0], C[j][1], af, bf);
                }
              }
            }
          }
          // epilogue: contract the product tile with phi^t'_b(r) for every t'
#pragma unroll
          for (int j = 0; j < NTW; j++) {
            const int nt = cg + CG * j;
            if (nt * 4 < nbc4) {
              const int bl = nt * 4 + lc;
#pragma unroll
              for (int t2 = 0; t2 < NT; t2++) {
                const double ph = sm.b[t2][bl][rh * 8 + lr];
                acc[s][sp][t2][0] += C[j][0] * ph;
                acc[s][sp][t2][1] += C[j][1] * ph;
              }
            }
          }
        }
      }
    }
  }
  // reduce over the 4 lanes of a row, then over column-group warps (fixed order: deterministic)
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem_raw);  // [8 warps][8 rows][2*2*NT*2]
  constexpr int NACC = 2 * 2 * NT * 2;
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int sp = 0; sp < 2; sp++)
#pragma unroll
      for (int t2 = 0; t2 < NT; t2++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double v = acc[s][sp][t2][c];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (lc == 0) red[((size_t)warp * 8 + lr) * NACC + ((s * 2 + sp) * NT + t2) * 2 + c] = v;
        }
  __syncthreads();
  const int ndd = NT * NT * 8;
  double* __restrict__ out = (is_kappa ? g.dd_kap : g.dd_rho) + ((size_t)za * 2 + q) * ndd * B.nghl;
  for (int idx = threadIdx.x; idx < NT * RT * NACC; idx += 256) {
    const int e = idx % NACC, rr = (idx / NACC) % RT, t = idx / (NACC * RT);
    const int c = e & 1, t2 = (e >> 1) % NT, ssp = e / (2 * NT);   // ssp = s*2+sp
    double v = 0.0;
    for (int k = 0; k < CG; k++) {
      const int w = (NT == 4) ? (t + 4 * (rr >> 3)) : ((rr >> 3) + 2 * k);
      v += red[((size_t)w * 8 + (rr & 7)) * NACC + e];
    }
    const int r = tile * RT + rr;
    if (r < B.nghl) out[(size_t)(((t * NT + t2) * 4 + ssp) * 2 + c) * B.nghl + r] = v;
  }
}

void launch_density(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  static bool attr = false;
  if (!attr) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(density_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DensSmem<4>)));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(density_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DensSmem<1>)));
    attr = true;
  }
  dim3 grid(a.basis.ntiles, 2, a.nactive);
  density_kernel<4><<<grid, 256, sizeof(DensSmem<4>), stream>>>(a, 0);
  density_kernel<1><<<grid, 256, sizeof(DensSmem<1>), stream>>>(a, 1);
}

// ================================================================================================
// pointwise fields: D -> 28 local densities -> mf(ta,tb,sa,sb)(r), pairing field pf(sa,sb)(r)
// ================================================================================================
// spin index convention: 0 = up (+1), 1 = down (-1)
__global__ void __launch_bounds__(128) fields_kernel(HamArgs g) {
  const DevBasis& B = g.basis;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y, za = blockIdx.z;
  if (r >= B.nghl) return;
  const size_t Ng = B.nghl;
  const double* __restrict__ dd = g.dd_rho + ((size_t)za * 2 + q) * NDD_RHO * Ng + r;
  const double* __restrict__ dk = g.dd_kap + ((size_t)za * 2 + q) * NDD_KAP * Ng + r;
  auto D = [&](int t, int t2, int s, int sp) -> cplx {
    const size_t e = (size_t)(((t * 4 + t2) * 4 + s * 2 + sp) * 2);
    return {dd[e * Ng], dd[(e + 1) * Ng]};
  };
  auto K = [&](int s, int sp) -> cplx {
    const size_t e = (size_t)((s * 2 + sp) * 2);
    return {dk[e * Ng], dk[(e + 1) * Ng]};
  };
  const cplx Z = {0.0, 0.0};
  cplx rho = Z, tau = Z, tjrr = Z, tjpr = Z, tjzr = Z, tjrp = Z, tjpp = Z, tjzp = Z, tjrz = Z, tjpz = Z, tjzz = Z;
  cplx sr = Z, sp_ = Z, sz = Z, tr = Z, tp = Z, tz = Z, jr = Z, jp = Z, jz = Z, fr = Z, fp = Z, fz = Z, gs = Z;
  cplx rb = Z, sbr = Z, sbp = Z, sbz = Z;
#pragma unroll
  for (int si = 0; si < 2; si++) {          // spin of |b>
    const double sg = si == 0 ? 1.0 : -1.0;
    const int so = 1 - si;                  // opposite spin (spin of |a> in the off-diagonal terms)
    // ---- diagonal in spin (pnfam_hamiltonian_blas.f90:263-357)
    cplx x = D(0, 0, si, si);
    rho = rho + x; sz = sz + sg * x;
    x = 0.5 * (D(0, 2, si, si) + D(2, 0, si, si));
    jp = jp + x; tjpz = tjpz + sg * x;
    x = D(1, 1, si, si) + D(2, 2, si, si) + D(3, 3, si, si);
    tau = tau + x; tz = tz + sg * x;
    x = mul_i(0.5 * (D(0, 1, si, si) - D(1, 0, si, si)));
    jr = jr + x; tjrz = tjrz + sg * x;
    x = mul_i(0.5 * (D(0, 3, si, si) - D(3, 0, si, si)));
    jz = jz + x; tjzz = tjzz + sg * x;
    gs = gs + sg * (D(0, 3, si, si) + D(3, 0, si, si));
    fr = fr + (0.5 * sg) * (D(1, 3, si, si) + D(3, 1, si, si));
    fp = fp + (0.5 * sg) * mul_i(D(2, 3, si, si) - D(3, 2, si, si));
    fz = fz + sg * D(3, 3, si, si);
    // ---- off-diagonal in spin (:360-572): |a> has spin -sg, |b> has spin sg
    x = D(0, 0, so, si);
    sr = sr + x; sp_ = sp_ + (-sg) * mul_i(x);
    x = 0.5 * (D(2, 0, so, si) + D(0, 2, so, si));
    tjpr = tjpr + x; tjpp = tjpp + (-sg) * mul_i(x);
    gs = gs + (-sg) * (D(0, 2, so, si) - D(2, 0, so, si));
    x = D(1, 1, so, si) + D(2, 2, so, si) + D(3, 3, so, si);
    tr = tr + x; tp = tp + (-sg) * mul_i(x);
    x = 0.5 * (D(0, 1, so, si) - D(1, 0, so, si));
    tjrr = tjrr + mul_i(x); tjrp = tjrp + sg * x;
    gs = gs + (D(0, 1, so, si) + D(1, 0, so, si));
    x = 0.5 * (D(0, 3, so, si) - D(3, 0, so, si));
    tjzr = tjzr + mul_i(x); tjzp = tjzp + sg * x;
    fr = fr + D(1, 1, so, si) + (-0.5 * sg) * (D(1, 2, so, si) - D(2, 1, so, si));
    fp = fp + mul_i((-sg) * D(2, 2, so, si) + 0.5 * (D(2, 1, so, si) - D(1, 2, so, si)));
    fz = fz + 0.5 * (D(3, 1, so, si) + D(1, 3, so, si) + (-sg) * (D(3, 2, so, si) - D(2, 3, so, si)));
    // ---- pairing densities (:640-675)
    x = 2.0 * K(so, si);
    rb = rb + (-sg) * x; sbz = sbz + x;
    x = 2.0 * K(si, si);
    sbr = sbr + (-sg) * x; sbp = sbp + mul_mi(x);
  }
  const double w = B.wdcori[r];
  rho = w * rho; tau = w * tau; tjrr = w * tjrr; tjpr = w * tjpr; tjzr = w * tjzr; tjrp = w * tjrp; tjpp = w * tjpp;
  tjzp = w * tjzp; tjrz = w * tjrz; tjpz = w * tjpz; tjzz = w * tjzz; sr = w * sr; sp_ = w * sp_; sz = w * sz;
  tr = w * tr; tp = w * tp; tz = w * tz; jr = w * jr; jp = w * jp; jz = w * jz; fr = w * fr; fp = w * fp; fz = w * fz;
  gs = w * gs; rb = w * rb; sbr = w * sbr; sbp = w * sbp; sbz = w * sbz;

  // ---- field tensor (pnfam_hamiltonian_blas.f90:775-1093), statement order preserved ------------
  const double crho = B.crho[r], cs = B.cs[r];
  const double ctau = B.ctau, cj = B.cj, ct = B.ct, cdrho = B.cdrho, cds = B.cds, crdj = B.crdj, csdj = B.csdj;
  const double ctj0 = B.ctj0, ctj1 = B.ctj1, ctj2 = B.ctj2, cf = B.cf, cgs = B.cgs;
  cplx mf[5][5][2][2];
#pragma unroll
  for (int a = 0; a < 5; a++)
#pragma unroll
    for (int b = 0; b < 5; b++)
#pragma unroll
      for (int s = 0; s < 4; s++) mf[a][b][s >> 1][s & 1] = Z;
  constexpr int P = 0, M = 1;
#define MF(a, b, s1, s2) mf[a][b][s1][s2]
#define ADD(a, b, s1, s2, sym, aux) MF(a, b, s1, s2) = MF(a, b, s1, s2) + (double)(sym) * (aux)
  const cplx t0 = ctj0 * (tjrr + tjpp + tjzz);
  const cplx t1_zr_rz = ctj1 * (tjzr - tjrz);
  const cplx t1_pz_zp = ctj1 * (tjpz - tjzp);
  const cplx t2_rz_zr = ctj2 * (tjrz + tjzr);
  const cplx t2_pz_zp = ctj2 * (tjpz + tjzp);
  cplx aux;
  // wf_a, wf_b, same spin
  MF(0, 0, P, P) = (2.0 * crho) * rho + ctau * tau;
  MF(0, 0, M, M) = MF(0, 0, P, P);
  aux = (2.0 * cs) * sz + ct * tz + cf * fz;
  ADD(0, 0, P, P, 1, aux); ADD(0, 0, M, M, -1, aux);
  // (0,1)/(1,0) same spin
  MF(0, 1, P, P) = (-crdj) * (tjpz - tjzp);
  MF(0, 1, M, M) = MF(0, 1, P, P);
  aux = (-csdj) * jp;
  ADD(0, 1, P, P, 1, aux); ADD(0, 1, M, M, -1, aux);
  MF(1, 0, P, P) = MF(0, 1, P, P);
  MF(1, 0, M, M) = MF(0, 1, M, M);
  aux = mul_i(t1_zr_rz) - 0.5 * mul_i(t2_rz_zr);
  ADD(0, 1, P, P, 1, aux); ADD(0, 1, M, M, -1, aux); ADD(1, 0, P, P, -1, aux); ADD(1, 0, M, M, 1, aux);
  aux = cj * mul_mi(jr);
  ADD(0, 1, P, P, 1, aux); ADD(0, 1, M, M, 1, aux); ADD(1, 0, P, P, -1, aux); ADD(1, 0, M, M, -1, aux);
  // (0,2)/(2,0) same spin
  MF(0, 2, P, P) = cj * jp;
  MF(0, 2, M, M) = MF(0, 2, P, P);
  aux = t1_pz_zp + 0.5 * t2_pz_zp;
  ADD(0, 2, P, P, 1, aux); ADD(0, 2, M, M, -1, aux);
  MF(2, 0, P, P) = MF(0, 2, P, P);
  MF(2, 0, M, M) = MF(0, 2, M, M);
  aux = csdj * mul_i(jr);
  ADD(0, 2, P, P, 1, aux); ADD(0, 2, M, M, -1, aux); ADD(2, 0, P, P, -1, aux); ADD(2, 0, M, M, 1, aux);
  aux = crdj * mul_mi(tjzr - tjrz);
  ADD(0, 2, P, P, 1, aux); ADD(0, 2, M, M, 1, aux); ADD(2, 0, P, P, -1, aux); ADD(2, 0, M, M, -1, aux);
  // (0,3)/(3,0) same spin
  MF(0, 3, P, P) = (-crdj) * (tjrp - tjpr);
  MF(0, 3, M, M) = MF(0, 3, P, P);
  aux = (2.0 * cgs) * gs;
  ADD(0, 3, P, P, 1, aux); ADD(0, 3, M, M, -1, aux);
  MF(3, 0, P, P) = MF(0, 3, P, P);
  MF(3, 0, M, M) = MF(0, 3, M, M);
  aux = mul_mi(t0) + (ctj2 / 3.0) * mul_i(tjrr + tjpp - 2.0 * tjzz);
  ADD(0, 3, P, P, 1, aux); ADD(0, 3, M, M, -1, aux); ADD(3, 0, P, P, -1, aux); ADD(3, 0, M, M, 1, aux);
  aux = cj * mul_mi(jz);
  ADD(0, 3, P, P, 1, aux); ADD(0, 3, M, M, 1, aux); ADD(3, 0, P, P, -1, aux); ADD(3, 0, M, M, -1, aux);
  // (0,4)/(4,0) same spin
  MF(0, 4, P, P) = (2.0 * cdrho) * rho;
  MF(0, 4, M, M) = MF(0, 4, P, P);
  aux = (2.0 * cds) * sz;
  ADD(0, 4, P, P, 1, aux); ADD(0, 4, M, M, -1, aux);
  MF(4, 0, P, P) = MF(0, 4, P, P);
  MF(4, 0, M, M) = MF(0, 4, M, M);
  // (1,2)/(2,1) same spin
  MF(1, 2, P, P) = csdj * sz;
  MF(1, 2, M, M) = MF(1, 2, P, P);
  aux = crdj * rho;
  ADD(1, 2, P, P, 1, aux); ADD(1, 2, M, M, -1, aux);
  MF(2, 1, P, P) = MF(1, 2, P, P);
  MF(2, 1, M, M) = MF(1, 2, M, M);
  // (1,3)/(3,1) same spin
  MF(1, 3, P, P) = (0.5 * cf) * sr;
  MF(1, 3, M, M) = -MF(1, 3, P, P);
  MF(3, 1, P, P) = MF(1, 3, P, P);
  MF(3, 1, M, M) = MF(1, 3, M, M);
  aux = csdj * mul_i(sp_);
  ADD(1, 3, P, P, 1, aux); ADD(1, 3, M, M, 1, aux); ADD(3, 1, P, P, -1, aux); ADD(3, 1, M, M, -1, aux);
  // (2,3)/(3,2) same spin
  MF(2, 3, P, P) = (-csdj) * sr;
  MF(2, 3, M, M) = MF(2, 3, P, P);
  aux = (0.5 * cf) * mul_mi(sp_);
  ADD(2, 3, P, P, 1, aux); ADD(2, 3, M, M, -1, aux);
  MF(3, 2, P, P) = MF(2, 3, M, M);
  MF(3, 2, M, M) = MF(2, 3, P, P);
  // (1,1),(2,2),(3,3) same spin
  MF(1, 1, P, P) = (4.0 * cdrho + ctau) * rho;
  MF(1, 1, M, M) = MF(1, 1, P, P);
  aux = (4.0 * cds + ct) * sz;
  ADD(1, 1, P, P, 1, aux); ADD(1, 1, M, M, -1, aux);
  MF(2, 2, P, P) = MF(1, 1, P, P);
  MF(2, 2, M, M) = MF(1, 1, M, M);
  MF(3, 3, P, P) = MF(1, 1, P, P);
  MF(3, 3, M, M) = MF(1, 1, M, M);
  aux = cf * sz;
  ADD(3, 3, P, P, 1, aux); ADD(3, 3, M, M, -1, aux);
  // ---- opposite spin
  MF(0, 0, P, M) = (2.0 * cs) * sr + ct * tr + cf * fr;
  MF(0, 0, M, P) = MF(0, 0, P, M);
  aux = mul_mi((2.0 * cs) * sp_ + ct * tp + cf * fp);
  ADD(0, 0, P, M, 1, aux); ADD(0, 0, M, P, -1, aux);
  MF(0, 1, P, M) = (2.0 * cgs) * gs;
  MF(0, 1, M, P) = MF(0, 1, P, M);
  aux = csdj * mul_mi(jz);
  ADD(0, 1, P, M, 1, aux); ADD(0, 1, M, P, -1, aux);
  MF(1, 0, P, M) = MF(0, 1, P, M);
  MF(1, 0, M, P) = MF(0, 1, M, P);
  MF(0, 2, P, M) = MF(0, 1, P, M);
  MF(2, 0, M, P) = MF(1, 0, M, P);
  MF(2, 0, P, M) = -MF(0, 2, P, M);
  MF(0, 2, M, P) = -MF(2, 0, M, P);
  aux = (-ctj1) * (tjrp - tjpr);
  ADD(0, 1, P, M, 1, aux); ADD(0, 1, M, P, -1, aux); ADD(1, 0, P, M, -1, aux); ADD(1, 0, M, P, 1, aux);
  ADD(0, 2, P, M, 1, aux); ADD(0, 2, M, P, 1, aux); ADD(2, 0, P, M, 1, aux); ADD(2, 0, M, P, 1, aux);
  aux = mul_mi(t0);
  ADD(0, 1, P, M, 1, aux); ADD(0, 1, M, P, 1, aux); ADD(1, 0, P, M, -1, aux); ADD(1, 0, M, P, -1, aux);
  ADD(0, 2, P, M, 1, aux); ADD(0, 2, M, P, -1, aux); ADD(2, 0, P, M, 1, aux); ADD(2, 0, M, P, -1, aux);
  aux = (-0.5 * ctj2) * (tjrp + tjpr);
  ADD(0, 1, P, M, 1, aux); ADD(0, 1, M, P, -1, aux); ADD(1, 0, P, M, -1, aux); ADD(1, 0, M, P, 1, aux);
  ADD(0, 2, P, M, -1, aux); ADD(0, 2, M, P, -1, aux); ADD(2, 0, P, M, -1, aux); ADD(2, 0, M, P, -1, aux);
  aux = (ctj2 / 3.0) * mul_i(-2.0 * tjrr + tjpp + tjzz);
  ADD(0, 1, P, M, 1, aux); ADD(0, 1, M, P, 1, aux); ADD(1, 0, P, M, -1, aux); ADD(1, 0, M, P, -1, aux);
  aux = (ctj2 / 3.0) * mul_i(tjrr - 2.0 * tjpp + tjzz);
  ADD(0, 2, P, M, 1, aux); ADD(0, 2, M, P, -1, aux); ADD(2, 0, P, M, 1, aux); ADD(2, 0, M, P, -1, aux);
  // (0,3)/(3,0) opposite spin
  MF(0, 3, P, M) = csdj * jp;
  MF(0, 3, M, P) = MF(0, 3, P, M);
  aux = csdj * mul_i(jr);
  ADD(0, 3, P, M, 1, aux); ADD(0, 3, M, P, -1, aux);
  MF(3, 0, P, M) = MF(0, 3, P, M);
  MF(3, 0, M, P) = MF(0, 3, M, P);
  aux = mul_mi(t1_zr_rz) - 0.5 * mul_i(t2_rz_zr);
  ADD(0, 3, P, M, 1, aux); ADD(0, 3, M, P, 1, aux); ADD(3, 0, P, M, -1, aux); ADD(3, 0, M, P, -1, aux);
  aux = t1_pz_zp - 0.5 * t2_pz_zp;
  ADD(0, 3, P, M, 1, aux); ADD(0, 3, M, P, -1, aux); ADD(3, 0, P, M, -1, aux); ADD(3, 0, M, P, 1, aux);
  // (0,4)/(4,0) opposite spin
  MF(0, 4, P, M) = (2.0 * cds) * sr;
  MF(0, 4, M, P) = MF(0, 4, P, M);
  aux = (2.0 * cds) * mul_mi(sp_);
  ADD(0, 4, P, M, 1, aux); ADD(0, 4, M, P, -1, aux);
  MF(4, 0, P, M) = MF(0, 4, P, M);
  MF(4, 0, M, P) = MF(0, 4, M, P);
  // (1,2)/(2,1) opposite spin
  MF(1, 2, P, M) = (0.5 * cf) * mul_i(sp_);
  MF(1, 2, M, P) = MF(1, 2, P, M);
  aux = (0.5 * cf) * sr;
  ADD(1, 2, P, M, 1, aux); ADD(1, 2, M, P, -1, aux);
  MF(2, 1, P, M) = -MF(1, 2, P, M);
  MF(2, 1, M, P) = -MF(1, 2, M, P);
  // (1,3),(3,1),(2,3),(3,2) opposite spin
  MF(1, 3, P, M) = (0.5 * cf) * sz;
  MF(1, 3, M, P) = MF(1, 3, P, M);
  aux = crdj * rho;
  ADD(1, 3, P, M, 1, aux); ADD(1, 3, M, P, -1, aux);
  MF(3, 1, P, M) = MF(1, 3, M, P);
  MF(3, 1, M, P) = MF(1, 3, P, M);
  MF(3, 2, P, M) = MF(3, 1, P, M);
  MF(2, 3, M, P) = MF(1, 3, M, P);
  MF(2, 3, P, M) = -MF(1, 3, P, M);
  MF(3, 2, M, P) = -MF(3, 1, M, P);
  // (3,3),(1,1),(2,2) opposite spin
  MF(3, 3, P, M) = (ct + 4.0 * cds) * sr;
  MF(3, 3, M, P) = MF(3, 3, P, M);
  aux = (ct + 4.0 * cds) * mul_mi(sp_);
  ADD(3, 3, P, M, 1, aux); ADD(3, 3, M, P, -1, aux);
  MF(1, 1, P, M) = MF(3, 3, P, M);
  MF(1, 1, M, P) = MF(3, 3, M, P);
  aux = cf * sr;
  ADD(1, 1, P, M, 1, aux); ADD(1, 1, M, P, 1, aux);
  MF(2, 2, P, M) = MF(3, 3, P, M);
  MF(2, 2, M, P) = MF(3, 3, M, P);
  aux = cf * mul_mi(sp_);
  ADD(2, 2, P, M, 1, aux); ADD(2, 2, M, P, -1, aux);
#undef ADD
#undef MF
  double* __restrict__ mo = g.mf + ((size_t)za * 2 + q) * NMF * Ng + r;
#pragma unroll
  for (int a = 0; a < 5; a++)
#pragma unroll
    for (int b = 0; b < 5; b++)
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const size_t e = (size_t)(((a * 5 + b) * 4 + s) * 2);
        mo[e * Ng] = mf[a][b][s >> 1][s & 1].re;
        mo[(e + 1) * Ng] = mf[a][b][s >> 1][s & 1].im;
      }
  // ---- pairing field (pnfam_hamiltonian_blas.f90:1201-1208): index (sa, sb)
  const double cp = B.cpair[r], csp = B.cspair[r];
  cplx pfv[2][2];
  pfv[0][0] = csp * (sbr + mul_mi(sbp));             // |a>=+, |b>=+
  pfv[1][0] = (-cp) * rb - csp * sbz;                // |a>=-, |b>=+
  pfv[0][1] = cp * rb - csp * sbz;                   // |a>=+, |b>=-
  pfv[1][1] = csp * (-sbr + mul_mi(sbp));            // |a>=-, |b>=-
  double* __restrict__ po = g.pf + ((size_t)za * 2 + q) * NPF * Ng + r;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    po[(size_t)(s * 2) * Ng] = pfv[s >> 1][s & 1].re;
    po[(size_t)(s * 2 + 1) * Ng] = pfv[s >> 1][s & 1].im;
  }
}

void launch_fields(const HamArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  dim3 grid((a.basis.nghl + 127) / 128, 2, a.nactive);
  fields_kernel<<<grid, 128, 0, stream>>>(a);
}

// ================================================================================================
// projection: h_ab = 2 sum_{r,t} phi^t_a(r) G^t_{s_a s_b}(r,b),  G^t = sum_t' mf^{t t'} phi^t'_b
// ================================================================================================
constexpr int GS = 68;   // padded row stride of G (64 interleaved (b,c) columns)

template <int NT>
struct ProjSmem {
  double a[NT][AC][RS];        // phi^t_a(r) chunk
  double b[NT][BC][RS];        // phi^t'_b(r) chunk
  double g[NT][RT][GS];        // G^t(r, (b,c))
  double mf[NT][NT][2][2][RT]; // field tensor for (sa fixed): [t][t'][sb][c][r]
};

// tile descriptor: x = block row, y = a-chunk start inside the block, z = b-chunk start, w = ksplit index
template <int NT>
__global__ void __launch_bounds__(256) projection_kernel(HamArgs g, const int4* __restrict__ tiles, int tile_off, int ntiles_q,
                                                          int ksplit, int q, int is_delta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ProjSmem<NT>& sm = *reinterpret_cast<ProjSmem<NT>*>(smem_raw);
  const DevBasis& B = g.basis;
  const int4 td = tiles[tile_off + blockIdx.x];
  const int ksp = blockIdx.y, za = blockIdx.z;
  const int ix = td.x, a0 = td.y, b0 = td.z;
  const DevBlockStruct st = is_delta ? g.d_out[q] : g.h_out[q];
  const int iy = st.r2c[ix];
  const int di = B.db[ix], dj = B.db[iy], nui = B.nsu[ix], nuj = B.nsu[iy];
  const int ia = B.isstart[ix], ib = B.isstart[iy];
  // the a-chunk lies inside one spin segment; the b-chunk may straddle (handled per column)
  const int sa = a0 < nui ? 0 : 1;
  const int a_hi = sa == 0 ? nui : di;
  const int nac = min(AC, a_hi - a0), nbc = min(BC, dj - b0);
  const int nac8 = (nac + 7) & ~7, nbc4 = (nbc + 3) & ~3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  const size_t Ng = B.nghl;
  const double* __restrict__ mfg = (is_delta ? g.pf : g.mf) + ((size_t)za * 2 + q) * (is_delta ? NPF : NMF) * Ng;
  const int tiles_per = (B.ntiles + ksplit - 1) / ksplit;
  const int kt0 = ksp * tiles_per, kt1 = min(B.ntiles, kt0 + tiles_per);
  double C[8][2];
#pragma unroll
  for (int j = 0; j < 8; j++) C[j][0] = C[j][1] = 0.0;
  for (int kt = kt0; kt < kt1; kt++) {
    const double* __restrict__ phit = B.phi + (size_t)kt * NTYPE * B.dqp * RT;
    __syncthreads();
    for (int idx = threadIdx.x; idx < NT * nac8 * RT; idx += 256) {
      const int rr = idx & (RT - 1), al = (idx / RT) % nac8, t = idx / (RT * nac8);
      sm.a[t][al][rr] = al < nac ? phit[((size_t)t * B.dqp + ia + a0 + al) * RT + rr] : 0.0;
    }
    for (int idx = threadIdx.x; idx < NT * nbc4 * RT; idx += 256) {
      const int rr = idx & (RT - 1), bl = (idx / RT) % nbc4, t = idx / (RT * nbc4);
      sm.b[t][bl][rr] = bl < nbc ? phit[((size_t)t * B.dqp + ib + b0 + bl) * RT + rr] : 0.0;
    }
    for (int idx = threadIdx.x; idx < NT * NT * 4 * RT; idx += 256) {
      const int rr = idx & (RT - 1), c = (idx / RT) & 1, sb = (idx / (2 * RT)) & 1, tt = idx / (4 * RT);
      const int t = tt / NT, t2 = tt % NT;
      const int r = kt * RT + rr;
      double v = 0.0;
      if (r < (int)Ng) {
        const size_t e = is_delta ? (size_t)((sa * 2 + sb) * 2 + c) : (size_t)(((t * 5 + t2) * 4 + sa * 2 + sb) * 2 + c);
        v = mfg[e * Ng + r];
      }
      sm.mf[t][t2][sb][c][rr] = v;
    }
    __syncthreads();
    // G^t(r,(b,c)) = sum_t' mf[t][t'][sb(b)] * phi^t'_b(r)
    for (int idx = threadIdx.x; idx < NT * RT * nbc4; idx += 256) {
      const int bl = idx % nbc4, rr = (idx / nbc4) % RT, t = idx / (nbc4 * RT);
      const int sb = (b0 + bl) < nuj ? 0 : 1;
      double gr = 0.0, gi = 0.0;
#pragma unroll
      for (int t2 = 0; t2 < NT; t2++) {
        const double ph = sm.b[t2][bl][rr];
        gr += sm.mf[t][t2][sb][0][rr] * ph;
        gi += sm.mf[t][t2][sb][1][rr] * ph;
      }
      sm.g[t][rr][2 * bl] = gr;
      sm.g[t][rr][2 * bl + 1] = gi;
    }
    __syncthreads();
    if (warp * 8 < nac8) {
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int ks = 0; ks < RT / 4; ks++) {
          const double af = sm.a[t][warp * 8 + lr][ks * 4 + lc];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (j * 4 < nbc4) {
              const double bf = sm.g[t][ks * 4 + lc][j * 8 + lr];
              dmma884(C[j][0], C[j][1], af, bf);
            }
          }
        }
    }
  }
  // write the partial (factor 2 of the reference's dgemm alpha applied in the reduction)
  const size_t pstride = 2 * g.nxy;   // re | im
  double* __restrict__ part = g.hpart + (((size_t)za * 2 + q) * 2 + is_delta) * (size_t)ksplit * pstride + (size_t)ksp * pstride;
  const size_t off = st.r2m[ix];
  const int al = warp * 8 + lr;
  if (al < nac) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int bl = j * 4 + lc;
      if (bl < nbc) {
        const size_t e = off + (size_t)(a0 + al) + (size_t)(b0 + bl) * di;
        part[e] = C[j][0];
        part[g.nxy + e] = C[j][1];
      }
    }
  }
}

// sum the split-K partials (fixed order) and scale by 2
__global__ void projection_reduce_kernel(HamArgs g, int ksplit) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y >> 1, is_delta = blockIdx.y & 1, za = blockIdx.z;
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

size_t projection_partial_elems(const ProjPlan& pp, size_t nxy) { return (size_t)2 * 2 * pp.ksplit * 2 * nxy; }

void launch_projection(const HamArgs& a, const ProjPlan& pp, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  static bool attr = false;
  if (!attr) {
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(projection_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProjSmem<5>)));
    PNFAM_CUDA_CHECK(cudaFuncSetAttribute(projection_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProjSmem<1>)));
    attr = true;
  }
  for (int q = 0; q < 2; q++) {
    if (pp.ntiles_h[q] > 0) {
      dim3 grid(pp.ntiles_h[q], pp.ksplit, a.nactive);
      projection_kernel<5><<<grid, 256, sizeof(ProjSmem<5>), stream>>>(a, pp.tiles_h, pp.tile_off_h[q], pp.ntiles_h[q], pp.ksplit, q, 0);
    }
    if (pp.ntiles_d[q] > 0) {
      dim3 grid(pp.ntiles_d[q], pp.ksplit, a.nactive);
      projection_kernel<1><<<grid, 256, sizeof(ProjSmem<1>), stream>>>(a, pp.tiles_d, pp.tile_off_d[q], pp.ntiles_d[q], pp.ksplit, q, 1);
    }
  }
  dim3 gr((unsigned)((2 * a.nxy + 255) / 256), 4, a.nactive);
  projection_reduce_kernel<<<gr, 256, 0, stream>>>(a, pp.ksplit);
}

}  // namespace pnfam
