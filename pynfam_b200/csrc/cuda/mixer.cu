// Block (d) of the FAM iteration, device-resident:
//   * H20/F20 assembly + energy-denominator update   dR = G(omega) o (quench*dH + F) [o T]
//       (pnfam_solver.f90:169-183, matrix_2qp :510-544, complex_emult_bbm pnfam_type_bbm.f90:288-327)
//   * modified Broyden mixing (Johnson 1988) with history M, w0 = 0.01
//       (pnfam_broyden.f90:117-216).  The reference rebuilds the M x M Gram matrix every iteration
//       (O(M^2) dots); only the row of the newest difference vector changes, so it is updated
//       incrementally here: 2*iter_used dots + one fused axpy sweep per iteration, all HBM-streaming.
//   * strength and cross-term contraction  S = -(1/pi) F.dR   (pnfam_solver.f90:190-203)
#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

__device__ __forceinline__ size_t pack_offset(int c, int k, size_t nxy) {
  return (k < 2) ? ((size_t)c * 2 + k) * nxy : (4 + (size_t)c * 2 + (k - 2)) * nxy;
}

// ---- Greens function update ---------------------------------------------------------------------
__global__ void greens_kernel(MixArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y, za = blockIdx.z;
  if (i >= a.nxy || za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  double hre = a.fqp[(size_t)k * a.nxy + i], him = 0.0;
  if (a.quench != 0.0) {
    hre += a.quench * a.hqp[(((size_t)p * 2 + 0) * 4 + k) * a.nxy + i];
    him = a.quench * a.hqp[(((size_t)p * 2 + 1) * 4 + k) * a.nxy + i];
  }
  const double wre = a.omega[2 * p], wim = a.omega[2 * p + 1];
  const double sgn = (k & 1) ? 1.0 : -1.0;            // X,P: E - omega ; Y,Q: E + omega
  const double dre = a.esum[(size_t)k * a.nxy + i] + sgn * wre, dim = sgn * wim;
  const double inv = -1.0 / (dre * dre + dim * dim);  // G = -1/(d) = -conj(d)/|d|^2
  const double gre = dre * inv, gim = -dim * inv;
  double zre = gre * hre - gim * him, zim = gim * hre + gre * him;
  if (a.tfac) {
    const double t = a.tfac[(size_t)k * a.nxy + i];
    zre *= t; zim *= t;
  }
  double* vo = a.vout + (size_t)p * a.n;
  vo[pack_offset(0, k, a.nxy) + i] = zre;
  vo[pack_offset(1, k, a.nxy) + i] = zim;
}

void launch_greens(const MixArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  dim3 grid((unsigned)((a.nxy + 255) / 256), a.nvec / 2, a.nactive);
  greens_kernel<<<grid, 256, 0, stream>>>(a);
}

// ---- block reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; i++) s += sh[i];   // fixed order
  return s;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; i++) s = fmax(s, sh[i]);
  return s;
}

// Where a slot stands in the reference's mixing schedule (broyden_method, pnfam_broyden.f90:117-216) after `iter`
// completed iterations of its point: iteration 0 (or M <= 0) mixes linearly, later ones run the modified Broyden update
// on a ring of M history slots.
struct BroPos {
  bool broyden;
  int iter_used, ipos, inext;   // history vectors in use, 0-based slot of the newest difference, slot written next
};
__device__ __forceinline__ BroPos bro_pos(int iter, int M) {
  BroPos b{false, 0, -1, 0};
  if (M <= 0 || iter == 0) return b;
  b.broyden = true;
  b.iter_used = iter - 1 < M ? iter - 1 : M;
  b.ipos = iter - 1 - ((iter - 2) / M) * M - 1;   // Fortran integer arithmetic (truncation toward zero), 1-based -> 0-based
  b.inext = iter - ((iter - 1) / M) * M - 1;
  return b;
}

// amplitudes of freshly admitted points start from zero (pnfam_solver.f90:63-71 without storage)
__global__ void __launch_bounds__(256) reset_kernel(MixArgs a) {
  const int za = blockIdx.y;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  if (a.slot_iter[p] != 0) return;
  double* v = a.vin + (size_t)p * a.n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (size_t)gridDim.x * blockDim.x) v[e] = 0.0;
}
void launch_reset(const MixArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  reset_kernel<<<dim3(64, a.nactive), 256, 0, stream>>>(a);
}

// vout <- vout - vin ; si partial ; (iter>=2) raw difference vectors into slot ipos + partial |df|^2
__global__ void __launch_bounds__(256) bro_diff_kernel(MixArgs a) {
  __shared__ double sh[8];
  const int za = blockIdx.y;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za], iter = a.slot_iter[p];
  const BroPos bp = bro_pos(iter, a.Mmode);
  const int ipos = (bp.broyden && iter >= 2) ? bp.ipos : -1;
  double* vo = a.vout + (size_t)p * a.n;
  const double* vi = a.vin + (size_t)p * a.n;
  double* dfp = ipos >= 0 ? a.df + ((size_t)p * a.M + ipos) * a.n : nullptr;
  double* dvp = ipos >= 0 ? a.dv + ((size_t)p * a.M + ipos) * a.n : nullptr;
  double mx = 0.0, ss = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (size_t)gridDim.x * blockDim.x) {
    const double d = vo[e] - vi[e];
    vo[e] = d;
    mx = fmax(mx, fabs(d));
    if (dfp) {
      const double f = d - dfp[e];
      dfp[e] = f;
      dvp[e] = vi[e] - dvp[e];
      ss += f * f;
    }
  }
  mx = block_max(mx, sh);
  ss = block_sum(ss, sh);
  if (threadIdx.x == 0) {
    a.red[((size_t)p * a.nred + blockIdx.x) * 2] = mx;
    a.red[((size_t)p * a.nred + blockIdx.x) * 2 + 1] = ss;
  }
}

__global__ void bro_finalize_kernel(MixArgs a) {
  if ((int)blockIdx.x >= a.ctrl->nactive) return;
  const int p = a.active[blockIdx.x];
  if (threadIdx.x == 0) {
    double mx = 0.0, ss = 0.0;
    for (int i = 0; i < a.nred; i++) {
      mx = fmax(mx, a.red[((size_t)p * a.nred + i) * 2]);
      ss += a.red[((size_t)p * a.nred + i) * 2 + 1];
    }
    a.si[p] = mx;
    a.normi[p] = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
  }
}

// no mixing (M<0): vin = vout_new ; linear: vin += alpha*(vout_new - vin).  vout holds the difference.
__global__ void bro_linear_kernel(MixArgs a) {
  const int za = blockIdx.y;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  if (bro_pos(a.slot_iter[p], a.Mmode).broyden) return;
  const double alpha = (a.Mmode < 0) ? 1.0 : a.alpha;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  a.vin[(size_t)p * a.n + e] += alpha * a.vout[(size_t)p * a.n + e];
}

// dots of the (raw) newest difference vector and of vout with every stored df_i.  One CTA owns a slice of BRO_SLICE
// elements of one point: the slices of df_new and vout stay in registers while the CTA streams the same slice of every
// history vector once (8 independent 8-byte loads per thread in flight), so the history is read exactly once per
// iteration at HBM rate.  Slice partials are summed in a fixed order by bro_solve_kernel (deterministic).  The warp
// partials of BRO_CHUNK history vectors at a time are staged in shared memory (any history size).
constexpr int BRO_SLICE = 2048;
constexpr int BRO_CHUNK = 64;
__global__ void __launch_bounds__(256) bro_dots_kernel(MixArgs a) {
  __shared__ double sh[BRO_CHUNK][8][2];
  const int sl = blockIdx.x, za = blockIdx.y;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  const BroPos bp = bro_pos(a.slot_iter[p], a.Mmode);
  if (!bp.broyden || bp.iter_used == 0) return;
  const int ipos = bp.ipos, iter_used = bp.iter_used;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t e0 = (size_t)sl * BRO_SLICE + threadIdx.x;
  const double* __restrict__ dfp = a.df + (size_t)p * a.M * a.n;
  const double* __restrict__ dfn = dfp + (size_t)ipos * a.n;
  const double* __restrict__ vo = a.vout + (size_t)p * a.n;
  double fn[8], v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const size_t e = e0 + (size_t)j * 256;
    fn[j] = e < a.n ? dfn[e] : 0.0;
    v[j] = e < a.n ? vo[e] : 0.0;
  }
  for (int c0 = 0; c0 < iter_used; c0 += BRO_CHUNK) {
    const int c1 = min(iter_used, c0 + BRO_CHUNK);
    for (int i = c0; i < c1; i++) {
      const double* __restrict__ dfi = dfp + (size_t)i * a.n;
      double f[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const size_t e = e0 + (size_t)j * 256;
        f[j] = e < a.n ? dfi[e] : 0.0;
      }
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) { s1 += f[j] * fn[j]; s2 += f[j] * v[j]; }
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) { sh[i - c0][warp][0] = s1; sh[i - c0][warp][1] = s2; }
    }
    __syncthreads();
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < 8; w++) { s1 += sh[i - c0][w][0]; s2 += sh[i - c0][w][1]; }
      double* part = a.dotpart + (((size_t)p * a.M + i) * a.nslices + sl) * 2;
      part[0] = s1; part[1] = s2;
    }
    __syncthreads();
  }
}

// gamma = B^{-1} work, B = Gram + w0^2 on the diagonal (SPD): Cholesky by one CTA, in shared memory while the factor fits
// (BRO_SOLVE_SMEM_N history vectors), else in the global work space a.chol (the reference: DSYTRF/DSYTRI, any M)
constexpr int BRO_SOLVE_SMEM_N = 64;
__global__ void bro_solve_kernel(MixArgs a) {
  extern __shared__ double Lsh[];
  if ((int)blockIdx.x >= a.ctrl->nactive) return;
  const int p = a.active[blockIdx.x];
  const BroPos bp = bro_pos(a.slot_iter[p], a.Mmode);
  if (!bp.broyden || bp.iter_used == 0) return;
  const int ipos = bp.ipos;
  const int n = bp.iter_used, M = a.M;
  double* L = n <= BRO_SOLVE_SMEM_N ? Lsh : a.chol + (size_t)p * M * (M + 2);
  double* G = a.gram + (size_t)p * M * M;
  // finish the dots: sum the slice partials in a fixed order; df_ipos is still un-normalised in memory, so the dots
  // with it are scaled instead
  {
    const double nm = a.normi[p];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double* part = a.dotpart + ((size_t)p * M + i) * a.nslices * 2;
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < a.nslices; k++) { s1 += part[2 * k]; s2 += part[2 * k + 1]; }
      if (i == ipos) {
        G[(size_t)i * M + i] = 1.0 + a.w0 * a.w0;
        a.work[(size_t)p * M + i] = s2 * nm;
      } else {
        G[(size_t)i * M + ipos] = G[(size_t)ipos * M + i] = s1 * nm;
        a.work[(size_t)p * M + i] = s2;
      }
    }
  }
  __syncthreads();
  // Cholesky of B (right-looking, the trailing update spread over the CTA), then the two triangular solves
  const int ld = n + 1, tid = threadIdx.x, nt = blockDim.x;
  double* y = L + (size_t)n * ld;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    L[i * ld + j] = (i == j) ? (1.0 + a.w0 * a.w0) : G[(size_t)i * M + j];
  }
  for (int i = tid; i < n; i += nt) y[i] = a.work[(size_t)p * M + i];
  __syncthreads();
  for (int j = 0; j < n; j++) {
    if (tid == 0) L[j * ld + j] = sqrt(L[j * ld + j]);
    __syncthreads();
    const double d = L[j * ld + j];
    for (int i = j + 1 + tid; i < n; i += nt) L[i * ld + j] /= d;
    __syncthreads();
    const int m = n - j - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int ii = idx / m, kk = idx - ii * m;
      if (kk <= ii) L[(j + 1 + ii) * ld + j + 1 + kk] -= L[(j + 1 + ii) * ld + j] * L[(j + 1 + kk) * ld + j];
    }
    __syncthreads();
  }
  for (int i = 0; i < n; i++) {          // L y = work
    if (tid == 0) y[i] /= L[i * ld + i];
    __syncthreads();
    const double yi = y[i];
    for (int r = i + 1 + tid; r < n; r += nt) y[r] -= L[r * ld + i] * yi;
    __syncthreads();
  }
  for (int i = n - 1; i >= 0; i--) {     // L^T gamma = y
    if (tid == 0) y[i] /= L[i * ld + i];
    __syncthreads();
    const double yi = y[i];
    for (int r = tid; r < i; r += nt) y[r] -= L[i * ld + r] * yi;
    __syncthreads();
  }
  double* gm = a.gamma + (size_t)p * M;
  for (int i = tid; i < n; i += nt) gm[i] = y[i];
}

// curv = alpha*vout - sum_i gamma_i (dv_i + alpha df_i); store (vout, vin) in slot inext; vin += curv
__global__ void __launch_bounds__(256) bro_update_kernel(MixArgs a) {
  const int za = blockIdx.y;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  const BroPos bp = bro_pos(a.slot_iter[p], a.Mmode);
  if (!bp.broyden) return;
  const int ipos = bp.ipos, inext = bp.inext, iter_used = bp.iter_used;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  const double alpha = a.alpha;
  const double vo = a.vout[(size_t)p * a.n + e];
  const double vi = a.vin[(size_t)p * a.n + e];
  double* df = a.df + (size_t)p * a.M * a.n + e;
  double* dv = a.dv + (size_t)p * a.M * a.n + e;
  double* du = a.du + (size_t)p * a.M * a.n + e;
  const double nm = a.normi[p];
  const double* gm = a.gamma + (size_t)p * a.M;
  double curv = alpha * vo;
  // The combination dv_i + alpha df_i of every normalised history entry is kept as a third array, so this sweep -- the
  // largest HBM stream of the iteration -- reads one array per entry instead of two; only the newest entry (ipos) is
  // still raw: it is normalised here and its combination stored (the same expression, hence the same bits as before).
  for (int i = 0; i < iter_used; i++) {
    double t;
    if (i == ipos) {
      const double f = df[(size_t)i * a.n] * nm, v = dv[(size_t)i * a.n] * nm;
      t = v + alpha * f;
      if (ipos != inext) { df[(size_t)i * a.n] = f; dv[(size_t)i * a.n] = v; du[(size_t)i * a.n] = t; }
    } else {
      t = du[(size_t)i * a.n];
    }
    curv = curv - gm[i] * t;
  }
  df[(size_t)inext * a.n] = vo;
  dv[(size_t)inext * a.n] = vi;
  a.vin[(size_t)p * a.n + e] = vi + curv;
}

// Every slot is at its own iteration, so all stages are launched every lock step and each CTA looks up what its slot
// needs (linear mixing for a fresh point, the Broyden update otherwise).
void launch_broyden(const MixArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  const unsigned nblk = (unsigned)((a.n + 255) / 256);
  bro_diff_kernel<<<dim3(a.nred, a.nactive), 256, 0, stream>>>(a);
  bro_finalize_kernel<<<a.nactive, 32, 0, stream>>>(a);
  bro_linear_kernel<<<dim3(nblk, a.nactive), 256, 0, stream>>>(a);
  if (a.Mmode <= 0) return;
  bro_dots_kernel<<<dim3(a.nslices, a.nactive), 256, 0, stream>>>(a);
  const int ns = a.Mmode < BRO_SOLVE_SMEM_N ? a.Mmode : BRO_SOLVE_SMEM_N;
  const size_t sh = ((size_t)ns * (ns + 1) + ns) * sizeof(double);
  bro_solve_kernel<<<a.nactive, 128, sh, stream>>>(a);
  bro_update_kernel<<<dim3(nblk, a.nactive), 256, 0, stream>>>(a);
}
int broyden_launches(const MixArgs& a) { return a.Mmode <= 0 ? 3 : 6; }

// ---- strength function ----------------------------------------------------------------------------------
// S = -(1/pi) F . dR (contract_bbm, pnfam_solver.f90:196-203): STR_SPLIT CTAs per (operator, point) sum slices of the
// quadrants, the last stage adds the slices in a fixed order (deterministic)
constexpr int STR_SPLIT = 32;
__global__ void __launch_bounds__(256) strength_kernel(MixArgs a) {
  __shared__ double sh[8];
  const int k = blockIdx.x, za = blockIdx.y, sp = blockIdx.z;
  if (za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  const double* g = a.gqp + (size_t)k * 4 * a.nxy;
  const double* v = a.vin + (size_t)p * a.n;
  const int nq = a.nvec / 2;
  const size_t per = (a.nxy + STR_SPLIT - 1) / STR_SPLIT, e0 = sp * per, e1 = e0 + per < a.nxy ? e0 + per : a.nxy;
  double sre = 0.0, sim = 0.0;
  for (int q = 0; q < nq; q++) {
    const double* gq = g + (size_t)q * a.nxy;
    const double* vr = v + pack_offset(0, q, a.nxy);
    const double* vi = v + pack_offset(1, q, a.nxy);
    for (size_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
      const double gg = gq[e];
      sre += gg * vr[e];
      sim += gg * vi[e];
    }
  }
  sre = block_sum(sre, sh);
  sim = block_sum(sim, sh);
  if (threadIdx.x == 0) {
    double* part = a.strpart + (((size_t)za * a.nstr + k) * STR_SPLIT + sp) * 2;
    part[0] = sre; part[1] = sim;
  }
}
__global__ void strength_final_kernel(MixArgs a) {
  const int k = blockIdx.x, za = blockIdx.y;
  if (threadIdx.x != 0 || za >= a.ctrl->nactive) return;
  const int p = a.active[za];
  const double* part = a.strpart + ((size_t)za * a.nstr + k) * STR_SPLIT * 2;
  double sre = 0.0, sim = 0.0;
  for (int s = 0; s < STR_SPLIT; s++) { sre += part[2 * s]; sim += part[2 * s + 1]; }
  const double pi = 3.14159265358979323846264338327950288;
  a.strength[((size_t)p * a.nstr + k) * 2] = -sre / pi;
  a.strength[((size_t)p * a.nstr + k) * 2 + 1] = -sim / pi;
}
int broyden_slices(size_t n) { return (int)((n + BRO_SLICE - 1) / BRO_SLICE); }
size_t strength_partial_elems(int npoints, int nstr) { return (size_t)npoints * nstr * STR_SPLIT * 2; }

void launch_strength(const MixArgs& a, cudaStream_t stream) {
  if (a.nactive <= 0) return;
  strength_kernel<<<dim3(a.nstr, a.nactive, STR_SPLIT), 256, 0, stream>>>(a);
  strength_final_kernel<<<dim3(a.nstr, a.nactive), 32, 0, stream>>>(a);
}

// ---- retire / admit -------------------------------------------------------------------------------------
// One CTA.  Thread za handles active slot za: counts the iteration (ifam's iter, pnfam_solver.f90:114-209), publishes the
// state of its point and decides whether it is finished (si < eps: converged; max_iter reached: interrupted).  Thread 0
// then hands the free slots to the pending points in admission order and compacts the active list (stable, serial: the
// slot a point lands in never changes its result, the order is kept only for reproducible scheduling).
__global__ void __launch_bounds__(1024) batch_control_kernel(BatchArgs b) {
  __shared__ int done[1024];
  const int nact = b.ctrl->nactive, step = b.ctrl->step;
  for (int za = threadIdx.x; za < nact; za += blockDim.x) {
    const int s = b.active[za], p = b.slot_point[s];
    const int it = b.slot_iter[s] + 1;
    b.slot_iter[s] = it;
    const double si = b.si[s];
    b.out_iters[p] = it;
    b.out_si[p] = si;
    for (int k = 0; k < 2 * b.nstr; k++) b.out_strength[(size_t)p * 2 * b.nstr + k] = b.strength[(size_t)s * 2 * b.nstr + k];
    if (b.out_trace) {
      double* t = b.out_trace + ((size_t)p * (b.max_iter + 1) + it) * 4;
      t[0] = si; t[1] = b.strength[(size_t)s * 2 * b.nstr]; t[2] = b.strength[(size_t)s * 2 * b.nstr + 1]; t[3] = (double)step;
    }
    const bool conv = si < b.eps;
    if (conv) b.out_conv[p] = 1;
    done[za] = (conv || it >= b.max_iter) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0, next = b.ctrl->next_pending, ndone = b.ctrl->ndone;
    for (int za = 0; za < nact; za++) {
      const int s = b.active[za];
      if (done[za]) {
        ndone++;
        if (next >= b.npoints) continue;            // nothing pending: the slot leaves the batch
        const int p = b.order[next++];
        b.slot_point[s] = p;
        b.slot_iter[s] = 0;
        b.omega[2 * s] = b.omega_pt[2 * p];
        b.omega[2 * s + 1] = b.omega_pt[2 * p + 1];
      }
      b.active[n++] = s;
    }
    b.ctrl->nactive = n; b.ctrl->next_pending = next; b.ctrl->ndone = ndone; b.ctrl->step = step + 1;
  }
}
void launch_batch_control(const BatchArgs& b, cudaStream_t stream) {
  if (b.nslots > 1024) throw std::runtime_error("batch control: more than 1024 slots");
  batch_control_kernel<<<1, 1024, 0, stream>>>(b);
}

}  // namespace pnfam
