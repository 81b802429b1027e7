// Device context, batched FAM solve loop and the C ABI of libpnfam_b200.so (include/pnfam_b200.h, section 2).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../../include/pnfam_b200.h"
#include "../host/symbolic.hpp"
#include "device_common.cuh"
#include "kernels.cuh"

namespace pnfam {

// ---- small RAII device buffer -------------------------------------------------------------------------
// Stream-ordered allocations from the device's default memory pool, which is told to keep freed memory: a solve
// needs tens of GB of work space (Broyden history), and a fresh context / solve then reuses what the previous one
// returned instead of going back to the driver.  All allocations and frees are ordered on the legacy default
// stream, with which the (blocking) work stream of a context synchronises implicitly.
// Two pools per device: the default one holds the large per-solve arrays (Broyden history, amplitudes: always the same
// sequence of sizes, so a solve finds the blocks the previous one returned), a second one the small allocations (tables of
// a context, task lists of an operator) -- otherwise those would be carved out of the returned large blocks and the next
// solve would have to go back to the driver for tens of GB (0.4-0.7 s per solve, seen in the end-to-end timing).
constexpr size_t SMALL_ALLOC_BYTES = (size_t)32 << 20;
inline cudaMemPool_t small_pool_of_current_device() {
  static bool done[64] = {};
  static cudaMemPool_t small[64] = {};
  static std::mutex guard;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(guard);
  if (!done[dev & 63]) {
    unsigned long long keep = ~0ull;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&small[dev & 63], &props) == cudaSuccess) cudaMemPoolSetAttribute(small[dev & 63], cudaMemPoolAttrReleaseThreshold, &keep);
    else { small[dev & 63] = nullptr; cudaGetLastError(); }
    done[dev & 63] = true;
  }
  return small[dev & 63];
}
inline void keep_pool_memory() { (void)small_pool_of_current_device(); }
// bytes a solve may still allocate on the current device: free device memory + what the pool holds but does not use
inline size_t device_bytes_available() {
  size_t fr = 0, tot = 0;
  PNFAM_CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
  int dev = 0;
  cudaMemPool_t pool;
  unsigned long long reserved = 0, used = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
    fr += (size_t)(reserved - used);
  return fr;
}
struct Event {
  cudaEvent_t e = nullptr;
  Event() { PNFAM_CUDA_CHECK(cudaEventCreate(&e)); }
  ~Event() { if (e) cudaEventDestroy(e); }
  Event(const Event&) = delete;
  Event& operator=(const Event&) = delete;
};
template <class T>
struct PinnedBuf {
  T* p = nullptr;
  explicit PinnedBuf(size_t n) { PNFAM_CUDA_CHECK(cudaMallocHost(&p, n * sizeof(T))); std::memset(p, 0, n * sizeof(T)); }
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
};
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() { if (p) cudaFreeAsync(p, 0); p = nullptr; n = 0; }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) {
      cudaMemPool_t sp = small_pool_of_current_device();
      if (sp && count * sizeof(T) < SMALL_ALLOC_BYTES) PNFAM_CUDA_CHECK(cudaMallocFromPoolAsync(&p, count * sizeof(T), sp, 0));
      else PNFAM_CUDA_CHECK(cudaMallocAsync(&p, count * sizeof(T), 0));
    }
  }
  void upload(const std::vector<T>& h) {
    alloc(h.size());
    if (!h.empty()) PNFAM_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void zero() { if (p) PNFAM_CUDA_CHECK(cudaMemset(p, 0, n * sizeof(T))); }
};

struct DevStructBuf {
  DBuf<int> r2c, r2m;
  void upload(const BlockStruct& s) { r2c.upload(s.r2c); r2m.upload(s.r2m); }
  void upload(const std::vector<int>& c, const std::vector<int>& m) { r2c.upload(c); r2m.upload(m); }
  DevBlockStruct view() const { return {r2c.p, r2m.p}; }
};

// Re-layout of the reference's (nghl, dqp) wave-function tables into the padded, slot-interleaved, rotated tables
// of device_common.cuh.  One thread per (grid point of the r-tile, padded row).
__global__ void build_tables_kernel(const double* __restrict__ raw, size_t nraw, const int* __restrict__ p2s, int dqp_p, int nghl,
                                    double* __restrict__ phi5, double* __restrict__ phi4, double* __restrict__ phi0) {
  const int tile = blockIdx.x, pr = blockIdx.y * blockDim.y + threadIdx.y, rr = threadIdx.x;
  if (pr >= dqp_p) return;
  const int st = p2s[pr], r = tile * RT + rr;
  if (st < 0 || r >= nghl) return;                          // padding stays zero
  const int pos = (rr + phi_rot(pr)) & (RT - 1);
#pragma unroll
  for (int t = 0; t < NTYPE; t++) {
    const double v = raw[(size_t)t * nraw + (size_t)st * nghl + r];
    phi5[(((size_t)tile * dqp_p + pr) * NTYPE + t) * RT + pos] = v;
    if (t < 4) phi4[(((size_t)tile * dqp_p + pr) * 4 + t) * RT + pos] = v;
    if (t == 0) phi0[(((size_t)(tile >> 2) * dqp_p + pr) * 4 + (tile & 3)) * RT + pos] = v;
  }
}

}  // namespace pnfam

using namespace pnfam;

struct pnfam_b200_ctx {
  int device = 0;
  int nb = 0, dqp = 0, nghl = 0, ntiles = 0;
  size_t dmat = 0;
  std::vector<int> db, isstart, nsu, pstart;
  int dqp_p = 0;                  // rows of the spin-segment-padded index space (device_common.cuh)
  std::vector<double> Ep, En, qp_fp, qp_fn;
  bool use_diag = false;
  DBuf<int> d_db, d_isstart, d_nsu, d_pstart;
  DBuf<double> d_phi4, d_phi5, d_phi0, d_wdcori, d_crho, d_cs, d_cpair, d_cspair;
  DBuf<double> d_Up, d_Vp, d_Un, d_Vn;
  DevBasis basis{};
  // sum-factorised path (separable basis, hamiltonian_sf.cu)
  SfDev sf{};
  std::vector<int> seg_prow0, seg_n, seg_nslots;   // per (block, spin) segment: first padded row, states, n_z slots
  std::vector<int> h_zrow, h_p2l;                   // host copies of the per-padded-row tables (list building)
  bool sf2 = false;                                 // fully factorised kernels (hamiltonian_sf2.cu)
  DBuf<double> d_zt, d_rg, d_rgp, d_r0q, d_rgt;
  DBuf<int> d_zrow, d_p2l, d_slot, d_segtab;
  cudaStream_t stream = nullptr;
  std::unique_ptr<SideStreams> side;   // side streams + fork/join events of this context
  int64_t launches = 0;
  int64_t table_h2d_bytes = 0;   // bytes of basis tables uploaded by ctx_create
  std::shared_ptr<void> ham_ws;  // persistent work space of pnfam_b200_calc_hamiltonian (HamWorkspace below)
};

// Set-up of the sum-factorised path from the separable factors of the model (include/pnfam_b200.h).  Returns false
// (general-table kernels are used) when the factors are absent or outside the tile limits of the kernels; throws when
// they do not reproduce the full tables.
static bool setup_separable(pnfam_b200_ctx& c, const pnfam_b200_model& m) {
  if (m.ngh <= 0 || m.ngl <= 0 || !m.sep_zrow || !m.sep_z || !m.sep_r || getenv("PNFAM_B200_GENERAL_TABLES")) return false;
  const int ngh = m.ngh, ngl = m.ngl, nzr = m.sep_nzrows, N = c.dqp, nghl = c.nghl;
  if ((size_t)ngh * ngl != (size_t)nghl) throw std::runtime_error("model: ngh*ngl != nghl");
  const int mt = (ngh + 7) / 8, kih = (ngh + 3) & ~3;
  if (2 * mt > 12 || 8 * kih > SF_THREADS) return false;   // warp roles of the density, thread map of the G build
  int zs = std::max(8 * mt, kih);
  while (zs % 16 != 4) zs++;
  // the factors must reproduce the tables
  {
    const double* tab[5] = {m.wf, m.wfdr, m.wfdp, m.wfdz, m.wfd2_all};
    double scale[5] = {0, 0, 0, 0, 0}, err[5] = {0, 0, 0, 0, 0};
#pragma omp parallel for schedule(static)
    for (int a = 0; a < N; a++) {
      double sc[5] = {0, 0, 0, 0, 0}, er[5] = {0, 0, 0, 0, 0};
      const int z = m.sep_zrow[a];
      if (z < 0 || z >= nzr) { er[0] = 1e300; sc[0] = 1; }
      else
        for (int il = 0; il < ngl; il++) {
          double R[4];
          for (int j = 0; j < 4; j++) R[j] = m.sep_r[((size_t)j * N + a) * ngl + il];
          for (int ih = 0; ih < ngh; ih++) {
            const double z0 = m.sep_z[((size_t)0 * nzr + z) * ngh + ih], z1 = m.sep_z[((size_t)1 * nzr + z) * ngh + ih];
            const double z2 = m.sep_z[((size_t)2 * nzr + z) * ngh + ih];
            const double v[5] = {z0 * R[0], z0 * R[1], z0 * R[2], z1 * R[0], z2 * R[0] + z0 * R[3]};
            const size_t o = (size_t)a * nghl + ih + (size_t)il * ngh;
            for (int t = 0; t < 5; t++) { sc[t] = std::max(sc[t], std::fabs(tab[t][o])); er[t] = std::max(er[t], std::fabs(tab[t][o] - v[t])); }
          }
        }
#pragma omp critical
      for (int t = 0; t < 5; t++) { scale[t] = std::max(scale[t], sc[t]); err[t] = std::max(err[t], er[t]); }
    }
    for (int t = 0; t < 5; t++)
      if (err[t] > 1e-11 * std::max(scale[t], 1e-300))
        throw std::runtime_error("model: the separable factors (sep_z, sep_r) do not reproduce wave-function table " + std::to_string(t));
  }
  // z rows: even and odd n_z in two contiguous groups, so that the slots of a spin segment (n_z of one parity in a
  // regular basis) are consecutive rows (conflict-free fragment loads)
  std::vector<int> rmap(nzr);
  for (int z = 0; z < nzr; z++) rmap[z] = (z >> 1) + (z & 1) * ((nzr + 1) / 2);
  std::vector<double> zt((size_t)3 * nzr * zs, 0.0);
  for (int k = 0; k < 3; k++)
    for (int z = 0; z < nzr; z++)
      for (int ih = 0; ih < ngh; ih++) zt[((size_t)k * nzr + rmap[z]) * zs + ih] = m.sep_z[((size_t)k * nzr + z) * ngh + ih];
  auto pad4 = [](int x) { return (x + 3) & ~3; };
  const int nseg = 2 * c.nb;
  c.seg_prow0.assign(nseg, 0); c.seg_n.assign(nseg, 0); c.seg_nslots.assign(nseg, 0);
  std::vector<int> zrow(c.dqp_p, 0), p2l(c.dqp_p, -1), slot(c.dqp_p, 0), segtab((size_t)nseg * SF_SEGTAB, 0), p2s(c.dqp_p, -1);
  for (int ib = 0; ib < c.nb; ib++)
    for (int sp = 0; sp < 2; sp++) {
      const int seg = 2 * ib + sp;
      const int first = sp == 0 ? 0 : c.nsu[ib], n = sp == 0 ? c.nsu[ib] : c.db[ib] - c.nsu[ib];
      const int prow0 = c.pstart[ib] + (sp == 0 ? 0 : pad4(c.nsu[ib]));
      c.seg_prow0[seg] = prow0; c.seg_n[seg] = n;
      std::vector<int> loc(n);
      for (int i = 0; i < n; i++) loc[i] = first + i;
      std::stable_sort(loc.begin(), loc.end(), [&](int x, int y) { return rmap[m.sep_zrow[c.isstart[ib] + x]] < rmap[m.sep_zrow[c.isstart[ib] + y]]; });
      int* st = &segtab[(size_t)seg * SF_SEGTAB];
      int ns = 0;
      for (int i = 0; i < n; i++) {
        const int zr = rmap[m.sep_zrow[c.isstart[ib] + loc[i]]];
        if (i == 0 || zr != zrow[prow0 + i - 1]) {
          if (ns >= SF_KMAX) return false;
          st[ns] = i; st[17 + ns] = zr; ns++;
        }
        zrow[prow0 + i] = zr; p2l[prow0 + i] = loc[i]; slot[prow0 + i] = ns - 1; p2s[prow0 + i] = c.isstart[ib] + loc[i];
      }
      st[ns] = n; st[33] = ns; st[34] = n;
      c.seg_nslots[seg] = ns;
    }
  std::vector<double> rg((size_t)ngl * c.dqp_p * 4, 0.0);
  for (int il = 0; il < ngl; il++)
    for (int pr = 0; pr < c.dqp_p; pr++)
      if (p2s[pr] >= 0)
        for (int j = 0; j < 4; j++) rg[((size_t)il * c.dqp_p + pr) * 4 + j] = m.sep_r[((size_t)j * N + p2s[pr]) * ngl + il];
  const int nilp = (ngl + 1) / 2;
  std::vector<double> rgp((size_t)nilp * c.dqp_p * 8, 0.0);
  for (int ip = 0; ip < nilp; ip++)
    for (int pr = 0; pr < c.dqp_p; pr++)
      for (int k = 0; k < 2; k++) {
        const int il = std::min(2 * ip + k, ngl - 1);
        for (int j = 0; j < 4; j++) rgp[((size_t)ip * c.dqp_p + pr) * 8 + k * 4 + j] = rg[((size_t)il * c.dqp_p + pr) * 4 + j];
      }
  const int nilq = (ngl + 3) / 4;
  std::vector<double> r0q((size_t)nilq * c.dqp_p * 4, 0.0);
  for (int iq = 0; iq < nilq; iq++)
    for (int pr = 0; pr < c.dqp_p; pr++)
      for (int k = 0; k < 4; k++)
        if (4 * iq + k < ngl) r0q[((size_t)iq * c.dqp_p + pr) * 4 + k] = rg[((size_t)(4 * iq + k) * c.dqp_p + pr) * 4];
  c.d_r0q.upload(r0q);
  {
    std::vector<double> rgt((size_t)ngl * 4 * c.dqp_p, 0.0);
    for (int il = 0; il < ngl; il++)
      for (int pr = 0; pr < c.dqp_p; pr++)
        for (int j = 0; j < 4; j++) rgt[((size_t)il * 4 + j) * c.dqp_p + pr] = rg[((size_t)il * c.dqp_p + pr) * 4 + j];
    c.d_rgt.upload(rgt);
  }
  c.h_zrow = zrow; c.h_p2l = p2l;
  c.d_rgp.upload(rgp);
  c.d_zt.upload(zt); c.d_rg.upload(rg); c.d_zrow.upload(zrow); c.d_p2l.upload(p2l); c.d_slot.upload(slot); c.d_segtab.upload(segtab);
  SfDev& S = c.sf;
  S.enabled = 1; S.ngh = ngh; S.ngl = ngl; S.mt = mt; S.kih = kih; S.zs = zs; S.nzrows = nzr; S.dqp_p = c.dqp_p;
  S.zt = c.d_zt.p; S.rg = c.d_rg.p; S.rgp = c.d_rgp.p; S.r0q = c.d_r0q.p; S.rgt = c.d_rgt.p; S.zrow = c.d_zrow.p; S.p2l = c.d_p2l.p; S.slot = c.d_slot.p; S.segtab = c.d_segtab.p;
  S.na_max = 1; S.kpad_max = 4;
  for (int seg = 0; seg < nseg; seg++) { S.na_max = std::max(S.na_max, c.seg_n[seg]); S.kpad_max = std::max(S.kpad_max, pad4(c.seg_nslots[seg])); }
  if (S.na_max > 8 * 12) return S.enabled = 0, false;
  // column chunk of the density: the largest that fits the shared memory of an SM
  S.nbc_max = 32;
  if (sf_density_smem_bytes(S) > 227 * 1024) S.nbc_max = 16;
  if (sf_density_smem_bytes(S) > 227 * 1024 || sf_projection_smem_bytes(S) > 227 * 1024) return S.enabled = 0, false;
  // both-sided factorisation (default; PNFAM_B200_SF1 keeps the one-sided kernels of hamiltonian_sf.cu)
  c.sf2 = !getenv("PNFAM_B200_SF1") && sf2_smem_bytes(S) <= 227 * 1024;
  return true;
}

static void set_err(char* err, int errlen, const std::string& s) {
  if (err && errlen > 0) {
    std::strncpy(err, s.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
  }
}

static void require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    throw std::runtime_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                             "): the pnFAM iteration has no CPU fallback");
  if (device < 0 || device >= n) throw std::runtime_error("invalid CUDA device index");
  PNFAM_CUDA_CHECK(cudaSetDevice(device));
}
// Every entry point works on the device of its context and hands the caller's current device back on return.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    require_device(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

extern "C" int pnfam_b200_ctx_create(const pnfam_b200_model* m, int device, pnfam_b200_ctx** out, char* err, int errlen) {
  try {
    DeviceGuard dg(device);
    auto c = std::make_unique<pnfam_b200_ctx>();
    c->device = device;
    c->nb = m->nb; c->dqp = m->dqp; c->nghl = m->nghl;
    c->ntiles = (m->nghl + RT - 1) / RT;
    c->db.assign(m->db, m->db + m->nb);
    c->nsu.assign(m->num_spin_up, m->num_spin_up + m->nb);
    c->isstart.resize(m->nb);
    int a = 0;
    c->dmat = 0;
    for (int i = 0; i < m->nb; i++) { c->isstart[i] = a; a += c->db[i]; c->dmat += (size_t)c->db[i] * c->db[i]; }
    if (a != m->dqp) throw std::runtime_error("model: sum(db) != dqp");
    c->Ep.assign(m->Ep, m->Ep + m->dqp);
    c->En.assign(m->En, m->En + m->dqp);
    c->use_diag = m->qp_fp != nullptr && m->qp_fn != nullptr;
    if (c->use_diag) { c->qp_fp.assign(m->qp_fp, m->qp_fp + m->dqp); c->qp_fn.assign(m->qp_fn, m->qp_fn + m->dqp); }
    c->d_db.upload(c->db); c->d_isstart.upload(c->isstart); c->d_nsu.upload(c->nsu);
    // ---- padded index space; wave-function tables in the layouts the kernels copy linearly (device_common.cuh)
    {
      auto pad4 = [](int x) { return (x + 3) & ~3; };
      c->pstart.resize(m->nb);
      std::vector<int> p2s;                                  // padded row -> state (-1: zero padding row)
      for (int i = 0; i < m->nb; i++) {
        c->pstart[i] = (int)p2s.size();
        for (int l = 0; l < pad4(c->nsu[i]); l++) p2s.push_back(l < c->nsu[i] ? c->isstart[i] + l : -1);
        for (int l = 0; l < pad4(c->db[i] - c->nsu[i]); l++) p2s.push_back(l < c->db[i] - c->nsu[i] ? c->isstart[i] + c->nsu[i] + l : -1);
      }
      c->dqp_p = (int)p2s.size();
      c->d_pstart.upload(c->pstart);
      // separable basis: the sum-factorised kernels need no Ng x N table on the device at all
      if (!setup_separable(*c, *m)) {
        // the reference's (nghl, dqp) tables go up as they are; the re-layout runs on the device
        const double* tab[NTYPE] = {m->wf, m->wfdr, m->wfdp, m->wfdz, m->wfd2_all};
        const size_t nraw = (size_t)c->dqp * c->nghl;
        DBuf<double> raw;
        raw.alloc(NTYPE * nraw);
        for (int t = 0; t < NTYPE; t++) PNFAM_CUDA_CHECK(cudaMemcpy(raw.p + t * nraw, tab[t], nraw * sizeof(double), cudaMemcpyHostToDevice));
        DBuf<int> d_p2s;
        d_p2s.upload(p2s);
        const int nsuper = (c->ntiles + 3) / 4;
        const size_t tail = (size_t)8 * NTYPE * RT;              // chunk copies may run a few rows past the last block
        c->d_phi5.alloc((size_t)c->ntiles * c->dqp_p * NTYPE * RT + tail); c->d_phi4.alloc((size_t)c->ntiles * c->dqp_p * 4 * RT + tail);
        c->d_phi0.alloc((size_t)nsuper * c->dqp_p * 4 * RT + tail);
        c->d_phi5.zero(); c->d_phi4.zero(); c->d_phi0.zero();
        build_tables_kernel<<<dim3(c->ntiles, (c->dqp_p + 7) / 8), dim3(RT, 8)>>>(raw.p, nraw, d_p2s.p, c->dqp_p, c->nghl, c->d_phi5.p, c->d_phi4.p,
                                                                             c->d_phi0.p);
        PNFAM_CUDA_CHECK(cudaGetLastError());
        PNFAM_CUDA_CHECK(cudaDeviceSynchronize());
        c->table_h2d_bytes = (int64_t)NTYPE * nraw * 8;
      } else {
        c->table_h2d_bytes = (int64_t)(c->d_zt.n + c->d_rg.n + c->d_rgp.n + c->d_r0q.n + c->d_rgt.n) * 8 + (int64_t)(c->d_zrow.n + c->d_p2l.n + c->d_slot.n + c->d_segtab.n) * 4;
      }
    }
    auto up = [&](DBuf<double>& d, const double* p, size_t n) { d.upload(std::vector<double>(p, p + n)); };
    up(c->d_wdcori, m->wdcori, m->nghl); up(c->d_crho, m->crho, m->nghl); up(c->d_cs, m->cs, m->nghl);
    up(c->d_cpair, m->cpair, m->nghl); up(c->d_cspair, m->cspair, m->nghl);
    up(c->d_Up, m->Up, c->dmat); up(c->d_Vp, m->Vp, c->dmat); up(c->d_Un, m->Un, c->dmat); up(c->d_Vn, m->Vn, c->dmat);
    DevBasis& B = c->basis;
    B.nb = c->nb; B.dqp = c->dqp; B.nghl = c->nghl; B.ntiles = c->ntiles;
    B.db = c->d_db.p; B.isstart = c->d_isstart.p; B.nsu = c->d_nsu.p;
    B.pstart = c->d_pstart.p; B.dqp_p = c->dqp_p; B.phi4 = c->d_phi4.p; B.phi5 = c->d_phi5.p; B.phi0 = c->d_phi0.p;
    B.wdcori = c->d_wdcori.p; B.crho = c->d_crho.p; B.cs = c->d_cs.p; B.cpair = c->d_cpair.p; B.cspair = c->d_cspair.p;
    B.cdrho = m->cdrho; B.ctau = m->ctau; B.ctj0 = m->ctj0; B.ctj1 = m->ctj1; B.ctj2 = m->ctj2; B.crdj = m->crdj;
    B.cds = m->cds; B.ct = m->ct; B.cj = m->cj; B.cgs = m->cgs; B.cf = m->cf; B.csdj = m->csdj;
    PNFAM_CUDA_CHECK(cudaStreamCreate(&c->stream));
    c->side = std::make_unique<SideStreams>();
    *out = c.release();
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

extern "C" int pnfam_b200_ctx_separable(const pnfam_b200_ctx* c) { return c && c->sf.enabled ? 1 : 0; }

extern "C" int64_t pnfam_b200_ctx_h2d_bytes(const pnfam_b200_ctx* c) {
  if (!c) return 0;
  return c->table_h2d_bytes + (int64_t)(5 * (size_t)c->nghl + 4 * c->dmat) * 8 + (int64_t)(3 * c->nb + c->nb) * 4;
}

extern "C" void pnfam_b200_ctx_destroy(pnfam_b200_ctx* c) {
  if (!c) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(c->device);
  c->ham_ws.reset();
  c->side.reset();
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  if (prev >= 0) cudaSetDevice(prev);
}

// ---- per-operator device data ----------------------------------------------------------------------
namespace {

struct Sf2Lists {                 // device lists of the fully factorised path (hamiltonian_sf2.cu)
  DBuf<int4> cols[4];             // rho q0, rho q1, kappa q0, kappa q1
  DBuf<int> el_src[4], cptr[4];
  int nelem[4] = {0, 0, 0, 0};
  DBuf<int2> zrange[4];
  DBuf<int> order[4];
  DBuf<Sf2Task> tasks[2][2];      // [h / Delta][pass]
  int ntasks[2][2] = {{0, 0}, {0, 0}};
  DBuf<unsigned char> need[2][2];
  void fill(Sf2Dev& F, int nzr) const {
    F.enabled = 1; F.nzr = nzr;
    for (int k = 0; k < 4; k++) {
      F.cols[k] = cols[k].p; F.el_src[k] = el_src[k].p; F.cptr[k] = cptr[k].p; F.nelem[k] = nelem[k];
      F.zrange[k] = zrange[k].p; F.order[k] = order[k].p;
    }
    for (int m = 0; m < 2; m++)
      for (int q = 0; q < 2; q++) { F.tasks[m][q] = tasks[m][q].p; F.ntasks[m][q] = ntasks[m][q]; F.need[m][q] = need[m][q].p; }
  }
};

struct FusedPlanBuf {
  DBuf<FusedJob> jobs;
  DBuf<int4> ctas[3];
};

struct OperatorDev {
  OperatorPlan plan;
  FusedPlanBuf fwd_fused, bwd_fused;
  DBuf<DevTask> fwd_tasks, bwd_tasks;
  DBuf<int4> fwd_tiles1, fwd_tiles2, bwd_tiles1, bwd_tiles2;
  DevicePlan fwd, bwd;
  DevStructBuf sp[4], hsp[4];
  DBuf<int4> tiles_h, tiles_d;
  ProjPlan proj;
  DBuf<double> gqp, esum, tfac;
  DBuf<DensStep> dsteps[4];       // rho pass0, rho pass1, kappa pass0, kappa pass1
  int ndsteps[4] = {0, 0, 0, 0};
  size_t scratch_elems = 0;
  size_t pk_rho = 0, pk_kap = 0;  // doubles of the packed rho / kappa chunks per (point, pass)
  // sum-factorised path
  DBuf<SfDensStep> sf_steps[4];
  int sf_nsteps[4] = {0, 0, 0, 0};
  DBuf<SfProjTile> sf_tiles[2][2];
  int sf_ntiles[2][2] = {{0, 0}, {0, 0}};
  int sf_ksplit = 1;
  Sf2Lists sf2;                   // fully factorised path
};

// fully factorised path: sub-block lists of the density (one list per input structure, built from the same steps as the
// packed images) and (row, column-run) tasks of the radial projection (one list per output structure)
struct Sf2Elems {                 // host image of one element list of the fully factorised density
  std::vector<int4> cols;
  std::vector<int> src, cptr, order;
  std::vector<int2> zrange;
};
Sf2Elems build_sf2_elems(const pnfam_b200_ctx& c, const BlockStruct& st) {
  const int nzr = c.sf.nzrows, npair = nzr * nzr;
  struct Sub { int na, nb, src_off, src_ld, a_row0, b_row0; };
  std::vector<std::vector<Sub>> byk((size_t)4 * npair);
  for (int ix = 0; ix < c.nb; ix++) {
    const int iy = st.r2c[ix];
    if (iy < 0) continue;
    for (int sw = 0; sw < 4; sw++) {
      const int sa = 2 * ix + (sw >> 1), sb = 2 * iy + (sw & 1);
      const int na = c.seg_n[sa], nbs = c.seg_n[sb], pa = c.seg_prow0[sa], pb = c.seg_prow0[sb];
      int i0 = 0;
      while (i0 < na) {
        const int zra = c.h_zrow[pa + i0];
        int i1 = i0 + 1;
        while (i1 < na && c.h_zrow[pa + i1] == zra) i1++;
        int j0 = 0;
        while (j0 < nbs) {
          const int zrb = c.h_zrow[pb + j0];
          int j1 = j0 + 1;
          while (j1 < nbs && c.h_zrow[pb + j1] == zrb) j1++;
          byk[(size_t)sw * npair + (size_t)zra * nzr + zrb].push_back(Sub{i1 - i0, j1 - j0, st.r2m[ix], c.db[ix], pa + i0, pb + j0});
          j0 = j1;
        }
        i0 = i1;
      }
    }
  }
  Sf2Elems E;
  E.cptr.assign((size_t)4 * (npair + 1), 0);
  E.order.assign((size_t)4 * (npair + 1), 0);
  E.zrange.assign((size_t)4 * nzr, make_int2(0, 0));
  for (int sw = 0; sw < 4; sw++) {
    std::vector<std::pair<long, int>> w;
    for (int p = 0; p < npair; p++) {
      E.cptr[(size_t)sw * (npair + 1) + p] = (int)E.cols.size();
      long work = 0;
      for (const Sub& e : byk[(size_t)sw * npair + p]) {
        for (int b = 0; b < e.nb; b++) {
          E.cols.push_back(make_int4((int)E.src.size(), e.na, e.a_row0, e.b_row0 + b));
          for (int a = 0; a < e.na; a++) E.src.push_back(e.src_off + c.h_p2l[e.a_row0 + a] + c.h_p2l[e.b_row0 + b] * e.src_ld);
        }
        work += (long)e.na * e.nb + 3 * e.nb;
      }
      if (work > 0) {
        w.push_back({-work, p});
        int2& r = E.zrange[(size_t)sw * nzr + p / nzr];
        const int z2 = p % nzr;
        if (r.x >= r.y) r = make_int2(z2, z2 + 1);
        else { r.x = std::min(r.x, z2); r.y = std::max(r.y, z2 + 1); }
      }
    }
    E.cptr[(size_t)sw * (npair + 1) + npair] = (int)E.cols.size();
    std::sort(w.begin(), w.end());              // non-empty (zr, zr') entries by decreasing work
    int* o = &E.order[(size_t)sw * (npair + 1)];
    o[0] = (int)w.size();
    for (size_t k = 0; k < w.size(); k++) o[1 + k] = w[k].second;
  }
  return E;
}

void build_sf2_tasks(const pnfam_b200_ctx& c, const BlockStruct& st, DBuf<Sf2Task>& d_tasks, int& ntasks, DBuf<unsigned char>& d_need) {
  const int nzr = c.sf.nzrows, npair = nzr * nzr;
  std::vector<Sf2Task> tasks;
  std::vector<unsigned char> need((size_t)4 * npair, 0);
  for (int ix = 0; ix < c.nb; ix++) {
    const int iy = st.r2c[ix];
    if (iy < 0) continue;
    for (int s = 0; s < 4; s++) {
      const int sa = 2 * ix + (s >> 1), sb = 2 * iy + (s & 1);
      const int na = c.seg_n[sa], nbs = c.seg_n[sb];
      if (na == 0 || nbs == 0) continue;
      int j0 = 0;
      while (j0 < nbs) {
        const int pb0 = c.seg_prow0[sb] + j0, zrb = c.h_zrow[pb0];
        int j1 = j0 + 1;
        while (j1 < nbs && j1 - j0 < SF2_RUN && c.h_zrow[c.seg_prow0[sb] + j1] == zrb) j1++;
        for (int i = 0; i < na;) {
          Sf2Task t{};
          t.pa = c.seg_prow0[sa] + i; t.pb0 = pb0; t.nb = j1 - j0; t.out_base = st.r2m[ix]; t.ld = c.db[ix]; t.sasb = s;
          t.na = (i + 1 < na && c.h_zrow[t.pa + 1] == c.h_zrow[t.pa]) ? 2 : 1;      // rows of equal n_z are neighbours
          tasks.push_back(t);
          need[(size_t)s * npair + (size_t)c.h_zrow[t.pa] * nzr + zrb] = 1;
          i += t.na;
        }
        j0 = j1;
      }
    }
  }
  ntasks = (int)tasks.size();
  if (tasks.empty()) tasks.push_back(Sf2Task{});
  d_tasks.upload(tasks); d_need.upload(need);
}

// in[k]: input structures (rho q0, rho q1, kappa q0, kappa q1); out[m][q]: output structures
void build_sf2_lists(const pnfam_b200_ctx& c, const BlockStruct* in[4], const BlockStruct* out[2][2], Sf2Lists& L) {
  for (int k = 0; k < 4; k++) {
    Sf2Elems E = build_sf2_elems(c, *in[k]);
    L.nelem[k] = (int)E.src.size();
    if (E.src.empty()) { E.cols.push_back(make_int4(0, 0, 0, 0)); E.src.push_back(0); }
    L.cols[k].upload(E.cols); L.el_src[k].upload(E.src); L.cptr[k].upload(E.cptr); L.order[k].upload(E.order); L.zrange[k].upload(E.zrange);
  }
  for (int m = 0; m < 2; m++)
    for (int q = 0; q < 2; q++) build_sf2_tasks(c, *out[m][q], L.tasks[m][q], L.ntasks[m][q], L.need[m][q]);
}

// sum-factorised density: (s, s') sweeps over the blocks of a structure, columns in chunks of nbc_max
size_t upload_sf_density_steps(const pnfam_b200_ctx& c, const BlockStruct& st, DBuf<SfDensStep>& buf, int& n,
                               std::vector<SfDensStep>* host_copy = nullptr) {
  std::vector<SfDensStep> h;
  size_t img = 0;
  for (int sweep = 0; sweep < 4; sweep++) {
    const size_t first = h.size();
    for (int ix = 0; ix < c.nb; ix++) {
      const int iy = st.r2c[ix];
      if (iy < 0) continue;
      const int sa = 2 * ix + (sweep >> 1), sb = 2 * iy + (sweep & 1);
      if (c.seg_n[sa] == 0 || c.seg_n[sb] == 0) continue;
      const int nb4 = (c.seg_n[sb] + 3) & ~3;
      for (int b0 = 0; b0 < nb4; b0 += c.sf.nbc_max) {
        SfDensStep d{};
        d.seg_a = sa; d.a_row0 = c.seg_prow0[sa]; d.na = c.seg_n[sa]; d.nslots = c.seg_nslots[sa];
        d.b_row0 = c.seg_prow0[sb] + b0; d.nbc = std::min(c.sf.nbc_max, nb4 - b0);
        d.img_off = (int)img; d.sweep = sweep; d.flags = 0; d.rho_off = st.r2m[ix]; d.ld = c.db[ix];
        img += (size_t)d.na * 2 * d.nbc;
        h.push_back(d);
      }
    }
    if (h.size() > first) h.back().flags |= 1;
  }
  if (img > 0x7fffffffull) throw std::runtime_error("density: packed image offsets overflow");
  n = (int)h.size();
  if (host_copy) *host_copy = h;
  if (h.empty()) h.push_back(SfDensStep{});
  buf.upload(h);
  return img;
}

void upload_sf_proj_tiles(const pnfam_b200_ctx& c, const BlockStruct& st, DBuf<SfProjTile>& buf, int& n) {
  // pair tasks (4 padded columns each), grouped by spin combination, the long ones first inside a group, groups
  // padded to multiples of 8 (one CTA = 8 tasks with the same field tensor)
  std::vector<SfProjTile> h;
  for (int s = 0; s < 4; s++) {
    std::vector<SfProjTile> grp;
    for (int ix = 0; ix < c.nb; ix++) {
      const int iy = st.r2c[ix];
      if (iy < 0) continue;
      const int sa = 2 * ix + (s >> 1), sb = 2 * iy + (s & 1);
      if (c.seg_n[sa] == 0 || c.seg_n[sb] == 0) continue;
      const int nb4 = (c.seg_n[sb] + 3) & ~3;
      for (int b0 = 0; b0 < nb4; b0 += 4) {
        SfProjTile t{};
        t.seg_a = sa; t.a_row0 = c.seg_prow0[sa]; t.na = c.seg_n[sa]; t.nslots = c.seg_nslots[sa];
        t.b_row0 = c.seg_prow0[sb] + b0; t.nbc = 4; t.sa = s >> 1; t.sb = s & 1;
        t.out_off = st.r2m[ix]; t.ld = c.db[ix];
        grp.push_back(t);
      }
    }
    std::stable_sort(grp.begin(), grp.end(), [](const SfProjTile& x, const SfProjTile& y) {
      const long cx = (x.nslots > 8 ? 2000 : 1000) + x.na, cy = (y.nslots > 8 ? 2000 : 1000) + y.na;
      return cx > cy;
    });
    while (grp.size() % 8 != 0) {
      SfProjTile t{};
      t.sa = s >> 1; t.sb = s & 1;
      grp.push_back(t);
    }
    h.insert(h.end(), grp.begin(), grp.end());
  }
  n = (int)h.size();
  if (h.empty()) h.push_back(SfProjTile{});
  buf.upload(h);
}

// returns the number of doubles of the packed rho chunks of this step list
size_t upload_density_steps(const pnfam_b200_ctx& c, const BlockStruct& st, DBuf<DensStep>& buf, int& n) {
  size_t pk = 0;
  build_density_steps(c.nb, c.db.data(), c.pstart.data(), c.nsu.data(), st.r2c.data(), st.r2m.data(), nullptr, &n, &pk);
  std::vector<DensStep> h(std::max(n, 1));
  build_density_steps(c.nb, c.db.data(), c.pstart.data(), c.nsu.data(), st.r2c.data(), st.r2m.data(), h.data(), &n, &pk);
  buf.upload(h);
  return pk;
}

int sf_ksplit_for(const pnfam_b200_ctx& c, const OperatorDev& od, int nactive) {
  const int per = std::max(1, od.sf_ntiles[0][0] / 8 * std::max(1, nactive));
  return std::min(c.sf.ngl, std::max(1, (2 * 148 + per - 1) / per));
}

void flatten(const TransformPlan& tp, DBuf<DevTask>& dt, DBuf<int4>& d1, DBuf<int4>& d2, DevicePlan& out) {
  std::vector<DevTask> tasks;
  std::vector<int4> t1, t2;
  int toff = 0, maxd = 0;
  auto tiles_of = [](std::vector<int4>& v, int task, int term, int m, int n) {
    for (int i = 0; i < m; i += 32)
      for (int j = 0; j < n; j += 32) v.push_back(make_int4(task, term, i, j));
  };
  for (const BlockTask& b : tp.tasks) {
    DevTask d{};
    d.out_quad = b.out_quad; d.out_off = b.out_off; d.m = b.m; d.n = b.n; d.nterms = b.nterms;
    maxd = std::max(maxd, std::max(b.m, b.n));
    for (int t = 0; t < b.nterms; t++) {
      const TripleTerm& s = b.t[t];
      DevTerm& x = d.t[t];
      x.a_mat = s.a_mat; x.a_off = s.a_off; x.a_trans = s.a_trans;
      x.b_quad = s.b_quad; x.b_off = s.b_off; x.b_trans = s.b_trans;
      x.c_mat = s.c_mat; x.c_off = s.c_off; x.c_trans = s.c_trans;
      x.alpha_re = s.alpha_re; x.alpha_im = s.alpha_im;
      x.t_off = toff;
      toff += b.m * b.n;
      tiles_of(t1, (int)tasks.size(), t, b.m, b.n);
    }
    tiles_of(t2, (int)tasks.size(), 0, b.m, b.n);
    tasks.push_back(d);
  }
  // longest tiles first (inner dimension x filled area), so the tail of a launch is made of short ones
  auto work1 = [&](const int4& e) { const DevTask& k = tasks[e.x]; return (long)k.m * std::min(32, k.m - e.z) * std::min(32, k.n - e.w); };
  auto work2 = [&](const int4& e) { const DevTask& k = tasks[e.x]; return (long)k.n * k.nterms * std::min(32, k.m - e.z) * std::min(32, k.n - e.w); };
  std::stable_sort(t1.begin(), t1.end(), [&](const int4& a, const int4& b) { return work1(a) > work1(b); });
  std::stable_sort(t2.begin(), t2.end(), [&](const int4& a, const int4& b) { return work2(a) > work2(b); });
  dt.upload(tasks); d1.upload(t1); d2.upload(t2);
  out.tasks = dt.p; out.tiles1 = d1.p; out.tiles2 = d2.p; out.ntasks = (int)tasks.size();
  out.ntiles1 = (int)t1.size(); out.ntiles2 = (int)t2.size();
  out.max_dim = maxd; out.scratch_elems = (size_t)toff;
}

// Jobs of the fused transform kernel (transform.cu).  Per output block the terms are grouped by their C operand,
// out = sum_x (sum_{t in x} alpha_t A_t B_t) C_x, and two output blocks whose groups are made of the same A_t B_t products
// (up to one factor per group) share phase 1:  the 8 (16 with P,Q) triple products per block row of the reference's
// triprod_bbm become 6 (8) products.  Structures the kernel has no shape for (more than 2 groups or 2 terms per group,
// blocks wider than 176) keep the two-phase kernels.
bool build_fused_host(const TransformPlan& tp, std::vector<FusedJob>& jobs, std::vector<int4> (&sorted)[3]) {
  struct Grp { std::vector<int> terms; };
  jobs.clear();
  std::map<std::vector<long long>, int> open;            // phase-1 signature -> job with a free output slot
  auto ab_key = [](const TripleTerm& t) { return (((((long long)t.a_mat * 2 + t.a_trans) * 8 + t.b_quad) * 2 + t.b_trans) << 40) ^ ((long long)t.a_off << 20) ^ t.b_off; };
  for (const BlockTask& b : tp.tasks) {
    if (b.n > 176) return false;
    std::vector<Grp> groups;
    for (int t = 0; t < b.nterms; t++) {
      bool found = false;
      for (Grp& g : groups) {
        const TripleTerm& r = b.t[g.terms[0]];
        if (r.c_mat == b.t[t].c_mat && r.c_off == b.t[t].c_off && r.c_trans == b.t[t].c_trans) { g.terms.push_back(t); found = true; break; }
      }
      if (!found) groups.push_back(Grp{{t}});
    }
    if (groups.size() > 2) return false;
    for (const Grp& g : groups) if (g.terms.size() > 2) return false;
    // signature of phase 1: dimensions + the (A, B) products of every group
    std::vector<long long> sig = {b.m, b.n, (long long)groups.size()};
    for (const Grp& g : groups) {
      sig.push_back((long long)g.terms.size());
      for (int t : g.terms) { sig.push_back(ab_key(b.t[t])); sig.push_back(b.t[t].a_off); sig.push_back(b.t[t].b_off); }
    }
    auto fill_out = [&](FusedOut& fo, const double* bre, const double* bim) {
      fo.out_quad = b.out_quad; fo.out_off = b.out_off;
      for (size_t x = 0; x < groups.size(); x++) {
        const TripleTerm& r = b.t[groups[x].terms[0]];
        fo.c_mat[x] = r.c_mat; fo.c_off[x] = r.c_off; fo.c_trans[x] = r.c_trans;
        fo.beta_re[x] = bre[x]; fo.beta_im[x] = bim[x];
      }
    };
    auto it = open.find(sig);
    bool paired = false;
    if (it != open.end()) {
      // same products in the same order: this block is beta_x times the phase-1 sums of the job, if the ratio of the
      // coefficients is one number per group and flavour
      FusedJob& J = jobs[it->second];
      double bre[2] = {0, 0}, bim[2] = {0, 0};
      bool ok = true;
      for (size_t x = 0; x < groups.size() && ok; x++)
        for (size_t k = 0; k < groups[x].terms.size() && ok; k++) {
          const TripleTerm& tt = b.t[groups[x].terms[k]];
          const FusedTerm& ft = J.t[x][k];
          if (ft.alpha_re == 0.0 || ft.alpha_im == 0.0) { ok = false; break; }
          const double rr = tt.alpha_re / ft.alpha_re, ri = tt.alpha_im / ft.alpha_im;
          if (k == 0) { bre[x] = rr; bim[x] = ri; }
          else if (rr != bre[x] || ri != bim[x]) ok = false;
        }
      if (ok) {
        fill_out(J.o[1], bre, bim);
        J.nout = 2;
        open.erase(it);
        paired = true;
      }
    }
    if (!paired) {
      FusedJob J{};
      J.m = b.m; J.n = b.n; J.ngroups = (int)groups.size(); J.nout = 1;
      for (size_t x = 0; x < groups.size(); x++) {
        J.nterms[x] = (int)groups[x].terms.size();
        for (size_t k = 0; k < groups[x].terms.size(); k++) {
          const TripleTerm& tt = b.t[groups[x].terms[k]];
          FusedTerm& ft = J.t[x][k];
          ft.a_mat = tt.a_mat; ft.a_off = tt.a_off; ft.a_trans = tt.a_trans;
          ft.b_quad = tt.b_quad; ft.b_off = tt.b_off; ft.b_trans = tt.b_trans;
          ft.alpha_re = tt.alpha_re; ft.alpha_im = tt.alpha_im;
        }
      }
      const double one[2] = {1.0, 1.0};
      fill_out(J.o[0], one, one);
      open[sig] = (int)jobs.size();
      jobs.push_back(J);
    }
  }
  // CTAs: strips of <= 4 m-tiles, evenly cut; three size classes (register shapes of the kernel); heaviest first
  std::vector<int4> ctas[3];
  std::vector<double> work[3];
  for (size_t j = 0; j < jobs.size(); j++) {
    const FusedJob& J = jobs[j];
    const int cls = J.n <= 88 ? 0 : (J.n <= 128 ? 1 : 2);
    const int m8 = (J.m + 7) / 8, nstrip = (m8 + 3) / 4, per = (m8 + nstrip - 1) / nstrip;
    double w = 0;
    for (int x = 0; x < J.ngroups; x++) w += (double)J.nterms[x] * J.m * J.n + (double)J.nout * J.n * J.n;
    for (int t0 = 0; t0 < m8; t0 += per) {
      ctas[cls].push_back(make_int4((int)j, t0, std::min(per, m8 - t0), 0));
      work[cls].push_back(w * std::min(per, m8 - t0));
    }
  }
  for (int k = 0; k < 3; k++) {
    std::vector<int> idx(ctas[k].size());
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b2) { return work[k][a] > work[k][b2]; });
    sorted[k].clear();
    for (int i : idx) sorted[k].push_back(ctas[k][i]);
  }
  return true;
}

bool build_fused(const TransformPlan& tp, FusedPlanBuf& fb, DevicePlan& out) {
  if (getenv("PNFAM_B200_TRANSFORM_2PHASE")) return false;
  std::vector<FusedJob> jobs;
  std::vector<int4> sorted[3];
  if (!build_fused_host(tp, jobs, sorted)) return false;
  fb.jobs.upload(jobs);
  out.jobs = fb.jobs.p;
  for (int k = 0; k < 3; k++) {
    fb.ctas[k].upload(sorted[k]);
    out.fctas[k] = fb.ctas[k].p; out.nfctas[k] = (int)sorted[k].size();
  }
  out.fused = true;
  out.njobs = (int)jobs.size();
  if (getenv("PNFAM_B200_TRANSFORM_DEBUG")) {
    double p_old = 0, p_new = 0;
    for (const BlockTask& b : tp.tasks) p_old += (double)b.nterms * ((double)b.m * b.m * b.n + (double)b.m * b.n * b.n);
    int n2 = 0;
    for (const FusedJob& J : jobs) {
      n2 += J.nout == 2;
      for (int x = 0; x < J.ngroups; x++) p_new += (double)J.nterms[x] * J.m * J.m * J.n + (double)J.nout * J.m * J.n * J.n;
    }
    std::fprintf(stderr, "[transform] %zu tasks -> %zu fused jobs (%d with two outputs), CTAs %d/%d/%d, products x%.3f\n", tp.tasks.size(),
                 jobs.size(), n2, out.nfctas[0], out.nfctas[1], out.nfctas[2], p_new / p_old);
  }
  return true;
}

// 2-quasiparticle tables of matrix_2qp (pnfam_solver.f90:510-544): v[e] = b*f1_i + c*f2_j on a structure
std::vector<double> table_2qp(const pnfam_b200_ctx& c, const BlockStruct& st, size_t nxy, double b, double cc,
                              const std::vector<double>& f1, const std::vector<double>& f2, double e0) {
  std::vector<double> v(nxy, 0.0);
  if (!st.allocated) return v;
  for (int ibr = 0; ibr < c.nb; ibr++) {
    const int ibc = st.r2c[ibr];
    if (ibc < 0) continue;
    size_t ipt = st.r2m[ibr];
    for (int i2 = 0; i2 < c.db[ibc]; i2++)
      for (int i1 = 0; i1 < c.db[ibr]; i1++) v[ipt++] = b * f1[c.isstart[ibr] + i1] + cc * f2[c.isstart[ibc] + i2] + e0;
  }
  return v;
}

void build_proj_tiles(const pnfam_b200_ctx& c, const BlockStruct st[2], std::vector<int4>& tiles, int ntiles[2], int off[2]) {
  for (int q = 0; q < 2; q++) {
    off[q] = (int)tiles.size();
    for (int ix = 0; ix < c.nb; ix++) {
      const int iy = st[q].r2c[ix];
      if (iy < 0) continue;
      // chunks in the padded index space of the blocks: a-chunks stay inside one spin segment
      auto pad4 = [](int x) { return (x + 3) & ~3; };
      const int di = c.db[ix], dj = c.db[iy], nu = c.nsu[ix];
      const int pj = pad4(c.nsu[iy]) + pad4(dj - c.nsu[iy]);
      for (int s = 0; s < 2; s++) {
        const int lo = s == 0 ? 0 : pad4(nu), hi = s == 0 ? pad4(nu) : pad4(nu) + pad4(di - nu);
        for (int a0 = lo; a0 < hi; a0 += 48)
          for (int b0 = 0; b0 < pj; b0 += 32) tiles.push_back(make_int4(ix, a0, b0, 0));
      }
    }
    ntiles[q] = (int)tiles.size() - off[q];
  }
}

// split-K factors of the projection for a batch of `nslots` slots (sizes the partial sums)
void set_batch_slots(const pnfam_b200_ctx& c, OperatorDev& od, int nslots) {
  const int S = std::max(1, nslots);
  if (c.sf.enabled) {
    // split-K factor as a function of the slots still active (the batch shrinks when the queue has drained): at least
    // two waves of CTAs per launch.  proj.ksplit sizes the partials for the largest nactive x ksplit(nactive).
    od.sf_ksplit = sf_ksplit_for(c, od, S);
    int need = od.sf_ksplit * S;
    for (int n = 1; n <= S; n++) need = std::max(need, n * sf_ksplit_for(c, od, n));
    od.proj.ksplit = (need + S - 1) / S;
  } else {
    const int per = std::max(1, od.proj.ntiles_h[0] * S);
    od.proj.ksplit = std::min(std::min(c.ntiles, 32), std::max(1, (2 * 148 + per - 1) / per));
  }
}

std::unique_ptr<OperatorDev> make_operator(pnfam_b200_ctx& c, const pnfam_b200_operator& op) {
  auto od = std::make_unique<OperatorDev>();
  std::vector<int> ir2c(op.f_ir2c, op.f_ir2c + c.nb);
  od->plan = make_operator_plan(c.db, ir2c, c.use_diag, op.beta_minus != 0);
  flatten(od->plan.forward, od->fwd_tasks, od->fwd_tiles1, od->fwd_tiles2, od->fwd);
  flatten(od->plan.backward, od->bwd_tasks, od->bwd_tiles1, od->bwd_tiles2, od->bwd);
  build_fused(od->plan.forward, od->fwd_fused, od->fwd);
  build_fused(od->plan.backward, od->bwd_fused, od->bwd);
  od->scratch_elems = std::max(od->fwd.scratch_elems, od->bwd.scratch_elems);
  for (int k = 0; k < 4; k++) { od->sp[k].upload(od->plan.sp[k]); od->hsp[k].upload(od->plan.hsp[k]); }
  if (c.sf.enabled) {
    std::vector<SfDensStep> hs[4];
    od->pk_rho = std::max(upload_sf_density_steps(c, od->plan.sp[0], od->sf_steps[0], od->sf_nsteps[0], &hs[0]),
                          upload_sf_density_steps(c, od->plan.sp[3], od->sf_steps[1], od->sf_nsteps[1], &hs[1]));
    od->pk_kap = std::max(upload_sf_density_steps(c, od->plan.sp[1], od->sf_steps[2], od->sf_nsteps[2], &hs[2]),
                          upload_sf_density_steps(c, od->plan.sp[2], od->sf_steps[3], od->sf_nsteps[3], &hs[3]));
    if (c.sf2) {
      const BlockStruct* in[4] = {&od->plan.sp[0], &od->plan.sp[3], &od->plan.sp[1], &od->plan.sp[2]};
      const BlockStruct* out[2][2] = {{&od->plan.hsp[0], &od->plan.hsp[3]}, {&od->plan.hsp[1], &od->plan.hsp[2]}};
      build_sf2_lists(c, in, out, od->sf2);
      od->pk_rho = od->pk_kap = 2 * od->plan.nxy;      // packed sub-blocks: every element once
    }
    upload_sf_proj_tiles(c, od->plan.hsp[0], od->sf_tiles[0][0], od->sf_ntiles[0][0]);
    upload_sf_proj_tiles(c, od->plan.hsp[3], od->sf_tiles[0][1], od->sf_ntiles[0][1]);
    upload_sf_proj_tiles(c, od->plan.hsp[1], od->sf_tiles[1][0], od->sf_ntiles[1][0]);
    upload_sf_proj_tiles(c, od->plan.hsp[2], od->sf_tiles[1][1], od->sf_ntiles[1][1]);
    return od;
  }
  od->pk_rho = std::max(upload_density_steps(c, od->plan.sp[0], od->dsteps[0], od->ndsteps[0]),
                        upload_density_steps(c, od->plan.sp[3], od->dsteps[1], od->ndsteps[1]));
  od->pk_kap = std::max(upload_density_steps(c, od->plan.sp[1], od->dsteps[2], od->ndsteps[2]),
                        upload_density_steps(c, od->plan.sp[2], od->dsteps[3], od->ndsteps[3]));
  // projection output tiles: pass 0 -> (h_pn = hsp[0], Delta+ = hsp[1]); pass 1 -> (h_np = hsp[3], Delta- = hsp[2])
  {
    std::vector<int4> th, td;
    BlockStruct sh[2] = {od->plan.hsp[0], od->plan.hsp[3]}, sd[2] = {od->plan.hsp[1], od->plan.hsp[2]};
    build_proj_tiles(c, sh, th, od->proj.ntiles_h, od->proj.tile_off_h);
    build_proj_tiles(c, sd, td, od->proj.ntiles_d, od->proj.tile_off_d);
    od->tiles_h.upload(th); od->tiles_d.upload(td);
    od->proj.tiles_h = od->tiles_h.p; od->proj.tiles_d = od->tiles_d.p;
  }
  return od;
}

struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double s() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

HamArgs make_ham_args(const pnfam_b200_ctx& c, const OperatorDev& od) {
  HamArgs h{};
  h.basis = c.basis;
  h.rho_in[0] = od.sp[0].view(); h.kap_in[0] = od.sp[1].view(); h.rho_in[1] = od.sp[3].view(); h.kap_in[1] = od.sp[2].view();
  h.h_out[0] = od.hsp[0].view(); h.d_out[0] = od.hsp[1].view(); h.h_out[1] = od.hsp[3].view(); h.d_out[1] = od.hsp[2].view();
  h.rho_quad[0] = 0; h.kap_quad[0] = 1; h.rho_quad[1] = 3; h.kap_quad[1] = 2;
  h.nxy = od.plan.nxy;
  for (int q = 0; q < 2; q++) {
    h.steps_rho[q] = od.dsteps[q].p; h.nsteps_rho[q] = od.ndsteps[q];
    h.steps_kap[q] = od.dsteps[2 + q].p; h.nsteps_kap[q] = od.ndsteps[2 + q];
  }
  h.pk_stride_rho = od.pk_rho; h.pk_stride_kap = od.pk_kap;
  h.sf = c.sf;
  h.side = c.side.get();
  if (c.sf.enabled) {
    for (int k = 0; k < 4; k++) { h.sf.steps[k] = od.sf_steps[k].p; h.sf.nsteps[k] = od.sf_nsteps[k]; }
    for (int m = 0; m < 2; m++)
      for (int q = 0; q < 2; q++) { h.sf.tiles[m][q] = od.sf_tiles[m][q].p; h.sf.ntiles[m][q] = od.sf_ntiles[m][q]; }
    h.sf.ksplit = od.sf_ksplit;
    h.sf.pk_stride[0] = od.pk_rho; h.sf.pk_stride[1] = od.pk_kap;
    if (c.sf2) od.sf2.fill(h.sf2, c.sf.nzrows);
  }
  return h;
}

}  // namespace

extern "C" int pnfam_b200_solve(pnfam_b200_ctx* c, const pnfam_b200_operator* op, const pnfam_b200_solver_params* prm,
                                int32_t npoints, const double* omega_re, const double* omega_im, double* strength,
                                int32_t* iters, int32_t* conv, double* si_out, double* trace, pnfam_b200_stats* stats,
                                char* err, int errlen) {
  try {
    Timer wall;
    DeviceGuard dg(c->device);
    cudaStream_t st = c->stream;
    const int P = npoints;
    if (P <= 0) return 0;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (prm->max_iter <= 0) {                      // ifam's loop body never runs: amplitudes and strengths stay zero
      for (int p = 0; p < P; p++) { iters[p] = 0; conv[p] = 0; si_out[p] = 1.0; }
      std::fill(strength, strength + (size_t)P * (1 + op->nxterms) * 2, 0.0);
      return 0;
    }
    int64_t launches = 0, h2d = 0, d2h = 0;
    auto od = make_operator(*c, *op);
    const size_t nxy = od->plan.nxy;
    const int nvec = c->use_diag ? 8 : 4, nq = nvec / 2;
    const size_t n = (size_t)nvec * nxy;
    const int nstr = 1 + op->nxterms;
    const double quench = prm->quench_residual_int;
    const bool no_residual = std::fabs(quench) < 1e-10;
    const int M = no_residual ? -1 : prm->broyden_history_size;
    const int Malloc = std::max(M, 1);
    const bool bminus = op->beta_minus != 0;
    const bool sf = c->sf.enabled != 0;
    const int nred = 64;
    const int max_iter = std::max(1, (int)prm->max_iter);

    // ---- slots: the work space of one omega point each; points beyond the slot count wait in the admission queue --
    const size_t mf_e = sf ? sf_mf_elems(c->sf.ngl, c->sf.kih) : mf_elems(c->ntiles);
    const size_t pf_e = sf ? sf_pf_elems(c->sf.ngl, c->sf.kih) : pf_elems(c->ntiles);
    auto slot_doubles = [&]() {
      size_t d = 2 * n;                                                        // vin, vout
      if (M > 0) d += 3 * (size_t)Malloc * n;                                  // Broyden history: df, dv, dv + alpha df
      d += (size_t)Malloc * Malloc + 2 * (size_t)Malloc + (size_t)Malloc * broyden_slices(n) * 2;
      if (M > 64) d += (size_t)Malloc * (Malloc + 2);
      d += (size_t)nred * 2 + 4 + (size_t)nstr * 2 + strength_partial_elems(1, nstr);
      d += 3 * 8 * nxy;                                                        // rsp, hsp, hqp
      d += 2 * std::max<size_t>(od->scratch_elems, 1);
      d += (size_t)2 * NDD_RHO * c->nghl + (size_t)2 * NDD_KAP * c->nghl + 2 * mf_e + 2 * pf_e;
      d += 2 * std::max<size_t>(od->pk_rho, 1) + 2 * std::max<size_t>(od->pk_kap, 1);
      if (c->sf2) d += 2 * (sf2_kt_elems(0, c->sf.ngl, c->sf.nzrows) + sf2_kt_elems(1, c->sf.ngl, c->sf.nzrows));
      return d;
    };
    const bool need_hpart = !(sf && c->sf2);     // the fully factorised projection writes h directly (no split-K partials)
    int S = std::min(P, 1024);
    {
      int cap = prm->batch_slots > 0 ? prm->batch_slots : 0;
      if (const char* e = getenv("PNFAM_B200_SLOTS")) cap = atoi(e);
      if (cap <= 0) cap = 64;                                                  // measured plateau of the throughput
      S = std::min(S, cap);
      const double avail = 0.92 * (double)device_bytes_available();
      for (;;) {
        set_batch_slots(*c, *od, S);
        const double need = 8.0 * ((double)S * (double)slot_doubles() + (need_hpart ? (double)S * (double)projection_partial_elems(od->proj, nxy) : 0.0));
        if (need <= avail || S == 1) break;
        S = std::max(1, std::min(S - 1, (int)(S * avail / need)));
      }
    }

    // ---- static per-operator tables ---------------------------------------------------------------
    std::vector<double> Ea = bminus ? c->Ep : c->En, Eb = bminus ? c->En : c->Ep;
    {
      const double sa = bminus ? prm->energy_shift_prot : prm->energy_shift_neut;
      const double sb = bminus ? prm->energy_shift_neut : prm->energy_shift_prot;
      for (auto& e : Ea) e += sa;
      for (auto& e : Eb) e += sb;
    }
    {
      std::vector<double> es(4 * nxy, 0.0), tf;
      for (int k = 0; k < nq; k++) {
        auto v = table_2qp(*c, od->plan.qp[k], nxy, 1.0, k < 2 ? 1.0 : -1.0, Ea, Eb, 0.0);
        std::copy(v.begin(), v.end(), es.begin() + (size_t)k * nxy);
      }
      od->esum.upload(es);
      if (c->use_diag) {
        const std::vector<double>& fa = bminus ? c->qp_fp : c->qp_fn;
        const std::vector<double>& fb = bminus ? c->qp_fn : c->qp_fp;
        tf.assign(4 * nxy, 0.0);
        for (int k = 0; k < 4; k++) {
          auto v = k < 2 ? table_2qp(*c, od->plan.qp[k], nxy, -1.0, -1.0, fa, fb, 1.0)
                         : table_2qp(*c, od->plan.qp[k], nxy, -1.0, +1.0, fa, fb, 0.0);
          std::copy(v.begin(), v.end(), tf.begin() + (size_t)k * nxy);
        }
        od->tfac.upload(tf);
      }
      h2d += (int64_t)(es.size() + tf.size()) * 8;
    }

    // ---- per-slot state -----------------------------------------------------------------------------
    DBuf<double> vin, vout, df, dv, du, gram, work, gamma, dotpart, chol, red, d_si, d_normi, d_omega, d_str, d_strpart;
    DBuf<double> rsp, hsp, hqp, scratch, dd_rho, dd_kap, mf, pf, hpart, pk_rho, pk_kap;
    vin.alloc((size_t)S * n); vout.alloc((size_t)S * n);
    vin.zero(); vout.zero();
    // Broyden history: only slots that have been written are ever read (iter_used bounds every loop): no clearing
    if (M > 0) { df.alloc((size_t)S * Malloc * n); dv.alloc((size_t)S * Malloc * n); du.alloc((size_t)S * Malloc * n); }
    gram.alloc((size_t)S * Malloc * Malloc); work.alloc((size_t)S * Malloc); gamma.alloc((size_t)S * Malloc);
    gram.zero(); work.zero(); gamma.zero();
    dotpart.alloc((size_t)S * Malloc * broyden_slices(n) * 2);
    if (M > 64) chol.alloc((size_t)S * Malloc * (Malloc + 2));
    red.alloc((size_t)S * nred * 2); d_si.alloc(S); d_normi.alloc(S); d_omega.alloc((size_t)S * 2);
    d_str.alloc((size_t)S * nstr * 2); d_strpart.alloc(strength_partial_elems(S, nstr));
    rsp.alloc((size_t)S * 8 * nxy); hsp.alloc((size_t)S * 8 * nxy); hqp.alloc((size_t)S * 8 * nxy);
    rsp.zero(); hsp.zero(); hqp.zero();
    scratch.alloc((size_t)S * 2 * std::max<size_t>(od->scratch_elems, 1));
    dd_rho.alloc((size_t)S * 2 * NDD_RHO * c->nghl); dd_kap.alloc((size_t)S * 2 * NDD_KAP * c->nghl);
    // field tensors are tile-major (kernels.cuh); the padding grid points are zeroed once and never written
    mf.alloc((size_t)S * 2 * mf_e); pf.alloc((size_t)S * 2 * pf_e);
    mf.zero(); pf.zero();
    if (sf) { dd_rho.zero(); dd_kap.zero(); }             // (s, s') sweeps without any step are never written
    pk_rho.alloc((size_t)S * 2 * std::max<size_t>(od->pk_rho, 1)); pk_kap.alloc((size_t)S * 2 * std::max<size_t>(od->pk_kap, 1));
    if (need_hpart) hpart.alloc((size_t)S * projection_partial_elems(od->proj, nxy));
    DBuf<double> kt0, kt1;
    if (c->sf2 && sf) {
      kt0.alloc((size_t)S * 2 * sf2_kt_elems(0, c->sf.ngl, c->sf.nzrows));
      kt1.alloc((size_t)S * 2 * sf2_kt_elems(1, c->sf.ngl, c->sf.nzrows));
    }

    // ---- batch control: admission order (the points that need the most iterations -- small |Im omega| -- first, so
    //      that the batch drains with the short ones), slot tables, per-point results -----------------------------
    std::vector<int> order(P);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(omega_im[a]) < std::fabs(omega_im[b]); });
    DBuf<int> d_active, d_slot_iter, d_slot_point, d_order, d_iters, d_conv;
    DBuf<double> d_omega_pt, d_out_si, d_out_str, d_trace;
    DBuf<BatchCtrl> d_ctrl;
    const int tstride = (max_iter + 1) * 4;
    {
      std::vector<int> act(S), sp(S), zero_s(S, 0), zero_p(P, 0);
      std::vector<double> w(2 * (size_t)P), ws(2 * (size_t)S);
      for (int p = 0; p < P; p++) { w[2 * p] = omega_re[p]; w[2 * p + 1] = omega_im[p]; }
      for (int s = 0; s < S; s++) { act[s] = s; sp[s] = order[s]; ws[2 * s] = w[2 * order[s]]; ws[2 * s + 1] = w[2 * order[s] + 1]; }
      d_active.upload(act); d_slot_point.upload(sp); d_slot_iter.upload(zero_s); d_order.upload(order);
      d_iters.upload(zero_p); d_conv.upload(zero_p);
      d_omega_pt.upload(w); d_omega.upload(ws);
      d_out_si.upload(std::vector<double>(P, 1.0));
      d_out_str.alloc((size_t)P * nstr * 2); d_out_str.zero();
      if (trace) { d_trace.alloc((size_t)P * tstride); d_trace.zero(); }
      d_ctrl.upload(std::vector<BatchCtrl>(1, BatchCtrl{S, S, 0, 0}));
      h2d += (int64_t)w.size() * 8 + (int64_t)ws.size() * 8 + (int64_t)(3 * S + 3 * P) * 4;
    }

    TransformArgs ta{};
    if (bminus) { ta.W[0] = c->d_Up.p; ta.W[1] = c->d_Vp.p; ta.W[2] = c->d_Un.p; ta.W[3] = c->d_Vn.p; }
    else { ta.W[0] = c->d_Un.p; ta.W[1] = c->d_Vn.p; ta.W[2] = c->d_Up.p; ta.W[3] = c->d_Vp.p; }
    ta.scratch = scratch.p; ta.scratch_stride = std::max<size_t>(od->scratch_elems, 1);
    ta.nxy = nxy; ta.active = d_active.p; ta.ctrl = d_ctrl.p;

    // ---- F and cross-term fields -> quasiparticle basis (pnfam_solver.f90:368-382): backward transform of
    //      dHsp = [f 0; 0 0] (real flavour), one field at a time through slot 0.
    od->gqp.alloc((size_t)nstr * 4 * nxy);
    od->gqp.zero();
    for (int k = 0; k < nstr; k++) {
      const double* src = k == 0 ? op->f_elem : op->g_elem[k - 1];
      PNFAM_CUDA_CHECK(cudaMemsetAsync(hsp.p, 0, 8 * nxy * sizeof(double), st));
      PNFAM_CUDA_CHECK(cudaMemcpyAsync(hsp.p, src, nxy * sizeof(double), cudaMemcpyHostToDevice, st));
      h2d += (int64_t)nxy * 8;
      TransformArgs b = ta;
      b.in = hsp.p; b.in_pstride = 8 * nxy; b.in_pack = 0;
      b.out = hqp.p; b.out_pstride = 8 * nxy; b.out_pack = 0;
      launches += launch_transform(od->bwd, b, 1, st);
      PNFAM_CUDA_CHECK(cudaMemcpyAsync(od->gqp.p + (size_t)k * 4 * nxy, hqp.p, 4 * nxy * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    PNFAM_CUDA_CHECK(cudaMemsetAsync(hsp.p, 0, 8 * nxy * sizeof(double), st));
    PNFAM_CUDA_CHECK(cudaMemsetAsync(hqp.p, 0, 8 * nxy * sizeof(double), st));

    HamArgs ha = make_ham_args(*c, *od);
    ha.rsp = rsp.p; ha.hsp = hsp.p; ha.dd_rho = dd_rho.p; ha.dd_kap = dd_kap.p; ha.mf = mf.p; ha.pf = pf.p;
    ha.pk_rho = pk_rho.p; ha.pk_kap = pk_kap.p;
    ha.sf.pk[0] = pk_rho.p; ha.sf.pk[1] = pk_kap.p;
    ha.hpart = hpart.p; ha.active = d_active.p; ha.ctrl = d_ctrl.p;
    ha.sf2.kt[0] = kt0.p; ha.sf2.kt[1] = kt1.p;

    MixArgs ma{};
    ma.nvec = nvec; ma.nxy = nxy; ma.n = n; ma.M = Malloc; ma.Mmode = M; ma.alpha = (double)0.7f; ma.w0 = 0.01;
    ma.hqp = hqp.p; ma.fqp = od->gqp.p; ma.esum = od->esum.p; ma.tfac = c->use_diag ? od->tfac.p : nullptr;
    ma.omega = d_omega.p; ma.quench = no_residual ? 0.0 : quench;
    ma.vin = vin.p; ma.vout = vout.p; ma.df = df.p; ma.dv = dv.p; ma.du = du.p; ma.gram = gram.p; ma.work = work.p; ma.gamma = gamma.p;
    ma.dotpart = dotpart.p; ma.nslices = broyden_slices(n); ma.chol = chol.p;
    ma.red = red.p; ma.nred = nred; ma.si = d_si.p; ma.normi = d_normi.p; ma.gqp = od->gqp.p; ma.nstr = nstr;
    ma.strength = d_str.p; ma.strpart = d_strpart.p; ma.active = d_active.p; ma.ctrl = d_ctrl.p; ma.slot_iter = d_slot_iter.p;

    BatchArgs ba{};
    ba.ctrl = d_ctrl.p; ba.active = d_active.p; ba.slot_iter = d_slot_iter.p; ba.slot_point = d_slot_point.p; ba.order = d_order.p;
    ba.npoints = P; ba.nslots = S; ba.max_iter = max_iter; ba.nstr = nstr; ba.eps = prm->convergence_epsilon;
    ba.si = d_si.p; ba.strength = d_str.p; ba.omega = d_omega.p; ba.omega_pt = d_omega_pt.p;
    ba.out_iters = d_iters.p; ba.out_conv = d_conv.p; ba.out_si = d_out_si.p; ba.out_strength = d_out_str.p;
    ba.out_trace = trace ? d_trace.p : nullptr;

    // ---- the loop of ifam (pnfam_solver.f90:114-209), all active slots in lock step.  The host never drains the
    //      stream inside the loop: it runs LAG steps ahead of the device and reads the control block of step k (a copy in
    //      pinned memory, completion known from an event) while steps k+1 .. k+LAG are queued.  The grids are sized with
    //      that stale -- never too small: the active count does not grow -- number of active slots.
    constexpr int LAG = 2, RING = LAG + 2;
    struct StepRec { Event done, d0, d1, p0, p1; bool timed = false; };
    StepRec ring[RING];
    PinnedBuf<BatchCtrl> snap(RING);
    Event ev0, ev1;
    PNFAM_CUDA_CHECK(cudaEventRecord(ev0.e, st));
    double t_dens = 0, t_proj = 0;
    int64_t n_dens = 0, n_proj = 0;
    std::vector<float> step_ms;
    int nact = S, launched = 0, harvested = 0;
    bool finished = false;
    auto harvest = [&](int k) {
      StepRec& r = ring[k % RING];
      PNFAM_CUDA_CHECK(cudaEventSynchronize(r.done.e));
      float ms = 0;
      if (r.timed) {
        cudaEventElapsedTime(&ms, r.d0.e, r.d1.e); t_dens += ms * 1e-3;
        cudaEventElapsedTime(&ms, r.p0.e, r.p1.e); t_proj += ms * 1e-3;
      }
      cudaEventElapsedTime(&ms, k == 0 ? ev0.e : ring[(k - 1) % RING].done.e, r.done.e);
      step_ms.push_back(ms);
      const BatchCtrl& bc = snap.p[k % RING];
      nact = std::min(nact, bc.nactive);
      if (bc.nactive == 0) finished = true;
      d2h += (int64_t)sizeof(BatchCtrl);
    };
    const int64_t max_steps = (int64_t)max_iter * ((P + S - 1) / S + 1) + LAG + 1;   // cannot be reached; guards the loop
    while (!finished && launched < max_steps) {
      StepRec& r = ring[launched % RING];
      ha.nactive = nact; ma.nactive = nact;
      if (sf) ha.sf.ksplit = sf_ksplit_for(*c, *od, nact);
      launch_reset(ma, st);
      launches += 1;
      r.timed = !no_residual;
      if (!no_residual) {
        TransformArgs f = ta;
        f.in = vin.p; f.in_pstride = n; f.in_pack = 1;
        f.out = rsp.p; f.out_pstride = 8 * nxy; f.out_pack = 0;
        launches += launch_transform(od->fwd, f, nact, st);
        PNFAM_CUDA_CHECK(cudaEventRecord(r.d0.e, st));
        launch_density(ha, st);
        PNFAM_CUDA_CHECK(cudaEventRecord(r.d1.e, st));
        launch_fields(ha, st);
        PNFAM_CUDA_CHECK(cudaEventRecord(r.p0.e, st));
        launch_projection(ha, od->proj, st);
        PNFAM_CUDA_CHECK(cudaEventRecord(r.p1.e, st));
        TransformArgs b = ta;
        b.in = hsp.p; b.in_pstride = 8 * nxy; b.in_pack = 0;
        b.out = hqp.p; b.out_pstride = 8 * nxy; b.out_pack = 0;
        launches += launch_transform(od->bwd, b, nact, st);
        launches += 3 + 1 + 5;   // pack + 2 densities, fields, 4 projections + reduce
        n_dens += 2; n_proj += 4;
      }
      launch_greens(ma, st);
      launch_broyden(ma, st);
      launch_strength(ma, st);
      launch_batch_control(ba, st);
      launches += 1 + broyden_launches(ma) + 2 + 1;
      PNFAM_CUDA_CHECK(cudaMemcpyAsync(&snap.p[launched % RING], d_ctrl.p, sizeof(BatchCtrl), cudaMemcpyDeviceToHost, st));
      PNFAM_CUDA_CHECK(cudaEventRecord(r.done.e, st));
      launched++;
      if (launched - harvested > LAG) harvest(harvested++);
    }
    while (harvested < launched) harvest(harvested++);
    if (!finished) throw std::runtime_error("solve: the batch did not drain (internal error)");
    PNFAM_CUDA_CHECK(cudaEventRecord(ev1.e, st));
    // ---- results ------------------------------------------------------------------------------------
    PNFAM_CUDA_CHECK(cudaMemcpyAsync(iters, d_iters.p, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, st));
    PNFAM_CUDA_CHECK(cudaMemcpyAsync(conv, d_conv.p, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, st));
    PNFAM_CUDA_CHECK(cudaMemcpyAsync(si_out, d_out_si.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
    PNFAM_CUDA_CHECK(cudaMemcpyAsync(strength, d_out_str.p, (size_t)P * nstr * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (trace) PNFAM_CUDA_CHECK(cudaMemcpyAsync(trace, d_trace.p, (size_t)P * tstride * sizeof(double), cudaMemcpyDeviceToHost, st));
    PNFAM_CUDA_CHECK(cudaStreamSynchronize(st));
    PNFAM_CUDA_CHECK(cudaGetLastError());
    d2h += (int64_t)P * (8 + 8 + nstr * 16) + (trace ? (int64_t)P * tstride * 8 : 0);
    int64_t total_iters = 0;
    for (int p = 0; p < P; p++) total_iters += iters[p];
    if (trace)
      for (int p = 0; p < P; p++) {
        double* t = trace + (size_t)p * tstride;
        t[0] = 1.0;
        for (int it = 1; it <= iters[p]; it++) {   // lock-step index -> seconds of that step
          const int k = (int)t[(size_t)it * 4 + 3];
          t[(size_t)it * 4 + 3] = (k >= 0 && k < (int)step_ms.size()) ? step_ms[k] * 1e-3 : 0.0;
        }
      }
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0.e, ev1.e);
    if (stats) {
      stats->seconds_total = wall.s(); stats->seconds_device = ms * 1e-3; stats->iterations = total_iters;
      stats->kernel_launches = launches; stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
      stats->seconds_density = t_dens; stats->seconds_projection = t_proj;
      stats->launches_density = n_dens; stats->launches_projection = n_proj;
      stats->flops_density = no_residual ? 0.0 : (double)total_iters * 2.0 * 20.0 * c->nghl * (double)nxy;
      stats->flops_projection = no_residual ? 0.0 : (double)total_iters * 2.0 * 24.0 * c->nghl * (double)nxy;
      stats->batch_slots = S; stats->lock_steps = launched;
    }
    c->launches += launches;
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

// ---- calc_hamiltonian-shaped entry -----------------------------------------------------------------------
// Work space of the plug-in entry, kept in the context between calls: the FAM loop of the reference calls
// calc_hamiltonian once per iteration with the SAME eight block structures (they are preset once per solve,
// pnfam_solver.f90:402-413), so step lists, tile lists and all device buffers are built on the first call and reused
// while the structures do not change; a call then costs the 8 + 8 element copies and the kernels.
namespace {
struct HamWorkspace {
  std::vector<int> key;           // the 8 + 8 (ir2c, ir2m) arrays and nelem the work space was built for
  size_t nxy = 0;
  DevStructBuf sin[4], sout[4];
  DBuf<double> rsp, hsp, dd_rho, dd_kap, mf, pf, hpart, pk_rho, pk_kap;
  DBuf<int> d_active;
  DBuf<DensStep> dsteps[4];
  DBuf<SfDensStep> sf_steps[4];
  DBuf<SfProjTile> sf_tiles[2][2];
  Sf2Lists sf2;
  DBuf<double> kt0, kt1;
  DBuf<int4> th, td;
  ProjPlan pp;
  HamArgs h{};
  PinnedBuf<double> stage;        // pinned staging of the 16 element arrays
  explicit HamWorkspace(size_t nxy_) : nxy(nxy_), stage(16 * nxy_) {}
};

std::vector<int> ham_key(const pnfam_b200_ctx& c, const pnfam_b200_blockmatrix in[8], const pnfam_b200_blockmatrix out[8]) {
  std::vector<int> k;
  k.reserve((size_t)16 * (2 * c.nb + 1));
  for (int io = 0; io < 2; io++)
    for (int i = 0; i < 8; i++) {
      const pnfam_b200_blockmatrix& b = io ? out[i] : in[i];
      k.push_back((int)b.nelem);
      k.insert(k.end(), b.ir2c, b.ir2c + c.nb);
      k.insert(k.end(), b.ir2m, b.ir2m + c.nb);
    }
  return k;
}

std::shared_ptr<HamWorkspace> build_ham_workspace(pnfam_b200_ctx* c, const pnfam_b200_blockmatrix in[8],
                                                  const pnfam_b200_blockmatrix out[8]) {
  size_t nxy = 0;
  for (int i = 0; i < 8; i++) nxy = std::max(nxy, (size_t)std::max(in[i].nelem, out[i].nelem));
  auto w = std::make_shared<HamWorkspace>(std::max<size_t>(nxy, 1));
  w->key = ham_key(*c, in, out);
  auto mk = [&](const pnfam_b200_blockmatrix& b, DevStructBuf& d, BlockStruct& hs) {
    hs.r2c.assign(c->nb, -1); hs.r2m.assign(c->nb, -1); hs.allocated = true;
    for (int i = 0; i < c->nb; i++) if (b.ir2c[i] > 0) { hs.r2c[i] = b.ir2c[i] - 1; hs.r2m[i] = b.ir2m[i] - 1; }
    d.upload(hs);
  };
  // argument order -> (pass, kind): in: rho_pn(0,1) k+(2,3) rho_np(4,5) k-(6,7); storage quads 0,1,3,2
  BlockStruct hin[4], hout[4];
  for (int pr = 0; pr < 4; pr++) { mk(in[2 * pr], w->sin[pr], hin[pr]); mk(out[2 * pr], w->sout[pr], hout[pr]); }
  w->rsp.alloc(8 * nxy); w->hsp.alloc(8 * nxy); w->rsp.zero(); w->hsp.zero();
  HamArgs& h = w->h;
  h.basis = c->basis;
  h.rho_in[0] = w->sin[0].view(); h.kap_in[0] = w->sin[1].view(); h.rho_in[1] = w->sin[2].view(); h.kap_in[1] = w->sin[3].view();
  h.h_out[0] = w->sout[0].view(); h.d_out[0] = w->sout[1].view(); h.h_out[1] = w->sout[2].view(); h.d_out[1] = w->sout[3].view();
  h.rho_quad[0] = 0; h.kap_quad[0] = 1; h.rho_quad[1] = 3; h.kap_quad[1] = 2;
  h.nxy = nxy;
  const bool sf = c->sf.enabled != 0;
  ProjPlan& pp = w->pp;
  h.sf = c->sf;
  h.side = c->side.get();
  h.ctrl = nullptr;
  if (sf) {
    // argument pairs: 0 rho_pn (pass 0), 1 kappa+ (pass 0), 2 rho_np (pass 1), 3 kappa- (pass 1)
    std::vector<SfDensStep> hs[4];
    h.pk_stride_rho = std::max(upload_sf_density_steps(*c, hin[0], w->sf_steps[0], h.sf.nsteps[0], &hs[0]), upload_sf_density_steps(*c, hin[2], w->sf_steps[1], h.sf.nsteps[1], &hs[1]));
    h.pk_stride_kap = std::max(upload_sf_density_steps(*c, hin[1], w->sf_steps[2], h.sf.nsteps[2], &hs[2]), upload_sf_density_steps(*c, hin[3], w->sf_steps[3], h.sf.nsteps[3], &hs[3]));
    if (c->sf2) {
      const BlockStruct* ins[4] = {&hin[0], &hin[2], &hin[1], &hin[3]};
      const BlockStruct* outs[2][2] = {{&hout[0], &hout[2]}, {&hout[1], &hout[3]}};
      build_sf2_lists(*c, ins, outs, w->sf2);
      h.pk_stride_rho = h.pk_stride_kap = 2 * nxy;
      h.sf.pk_stride[0] = h.sf.pk_stride[1] = 2 * nxy;
      w->sf2.fill(h.sf2, c->sf.nzrows);
      w->kt0.alloc(2 * sf2_kt_elems(0, c->sf.ngl, c->sf.nzrows)); w->kt1.alloc(2 * sf2_kt_elems(1, c->sf.ngl, c->sf.nzrows));
      h.sf2.kt[0] = w->kt0.p; h.sf2.kt[1] = w->kt1.p;
    }
    for (int k = 0; k < 4; k++) h.sf.steps[k] = w->sf_steps[k].p;
    upload_sf_proj_tiles(*c, hout[0], w->sf_tiles[0][0], h.sf.ntiles[0][0]);
    upload_sf_proj_tiles(*c, hout[2], w->sf_tiles[0][1], h.sf.ntiles[0][1]);
    upload_sf_proj_tiles(*c, hout[1], w->sf_tiles[1][0], h.sf.ntiles[1][0]);
    upload_sf_proj_tiles(*c, hout[3], w->sf_tiles[1][1], h.sf.ntiles[1][1]);
    for (int m = 0; m < 2; m++)
      for (int q = 0; q < 2; q++) h.sf.tiles[m][q] = w->sf_tiles[m][q].p;
    h.sf.ksplit = std::min(c->sf.ngl, std::max(1, (2 * 148 + std::max(1, h.sf.ntiles[0][0] / 8) - 1) / std::max(1, h.sf.ntiles[0][0] / 8)));
    pp.ksplit = h.sf.ksplit;
    h.sf.pk_stride[0] = h.pk_stride_rho; h.sf.pk_stride[1] = h.pk_stride_kap;
  } else {
    int ndsteps[4];
    h.pk_stride_rho = std::max(upload_density_steps(*c, hin[0], w->dsteps[0], ndsteps[0]), upload_density_steps(*c, hin[2], w->dsteps[1], ndsteps[1]));
    h.pk_stride_kap = std::max(upload_density_steps(*c, hin[1], w->dsteps[2], ndsteps[2]), upload_density_steps(*c, hin[3], w->dsteps[3], ndsteps[3]));
    for (int q = 0; q < 2; q++) {
      h.steps_rho[q] = w->dsteps[q].p; h.nsteps_rho[q] = ndsteps[q];
      h.steps_kap[q] = w->dsteps[2 + q].p; h.nsteps_kap[q] = ndsteps[2 + q];
    }
    std::vector<int4> vh, vd;
    BlockStruct shh[2] = {hout[0], hout[2]}, sdd[2] = {hout[1], hout[3]};
    build_proj_tiles(*c, shh, vh, pp.ntiles_h, pp.tile_off_h);
    build_proj_tiles(*c, sdd, vd, pp.ntiles_d, pp.tile_off_d);
    w->th.upload(vh); w->td.upload(vd);
    pp.tiles_h = w->th.p; pp.tiles_d = w->td.p;
    const int per = std::max(1, pp.ntiles_h[0]);
    pp.ksplit = std::min(std::min(c->ntiles, 32), std::max(1, (2 * 148 + per - 1) / per));
  }
  w->pk_rho.alloc(2 * std::max<size_t>(h.pk_stride_rho, 1)); w->pk_kap.alloc(2 * std::max<size_t>(h.pk_stride_kap, 1));
  h.pk_rho = w->pk_rho.p; h.pk_kap = w->pk_kap.p;
  h.sf.pk[0] = w->pk_rho.p; h.sf.pk[1] = w->pk_kap.p;
  w->dd_rho.alloc((size_t)2 * NDD_RHO * c->nghl); w->dd_kap.alloc((size_t)2 * NDD_KAP * c->nghl);
  w->dd_rho.zero(); w->dd_kap.zero();
  w->mf.alloc((size_t)2 * (sf ? sf_mf_elems(c->sf.ngl, c->sf.kih) : mf_elems(c->ntiles)));
  w->pf.alloc((size_t)2 * (sf ? sf_pf_elems(c->sf.ngl, c->sf.kih) : pf_elems(c->ntiles)));
  w->mf.zero(); w->pf.zero();
  w->hpart.alloc(projection_partial_elems(pp, nxy));
  w->d_active.upload(std::vector<int>{0});
  h.rsp = w->rsp.p; h.hsp = w->hsp.p; h.dd_rho = w->dd_rho.p; h.dd_kap = w->dd_kap.p; h.mf = w->mf.p; h.pf = w->pf.p; h.hpart = w->hpart.p;
  h.active = w->d_active.p; h.nactive = 1;
  return w;
}
}  // namespace

extern "C" int pnfam_b200_calc_hamiltonian(pnfam_b200_ctx* c, const pnfam_b200_blockmatrix in[8], pnfam_b200_blockmatrix out[8],
                                           char* err, int errlen) {
  try {
    DeviceGuard dg(c->device);
    cudaStream_t st = c->stream;
    auto w = std::static_pointer_cast<HamWorkspace>(c->ham_ws);
    if (!w || w->key != ham_key(*c, in, out)) {
      c->ham_ws.reset();
      w = build_ham_workspace(c, in, out);
      c->ham_ws = w;
    }
    const size_t nxy = w->h.nxy;
    const int quad_of_pair[4] = {0, 1, 3, 2};
    // elements: host arrays -> pinned staging -> device, all on the context's stream
    for (int pr = 0; pr < 4; pr++)
      for (int cc = 0; cc < 2; cc++) {
        const pnfam_b200_blockmatrix& b = in[2 * pr + cc];
        double* sg = w->stage.p + (size_t)(2 * pr + cc) * nxy;
        std::memcpy(sg, b.elem, (size_t)b.nelem * sizeof(double));
        PNFAM_CUDA_CHECK(cudaMemcpyAsync(w->rsp.p + ((size_t)cc * 4 + quad_of_pair[pr]) * nxy, sg, (size_t)b.nelem * sizeof(double),
                                         cudaMemcpyHostToDevice, st));
      }
    launch_density(w->h, st);
    launch_fields(w->h, st);
    launch_projection(w->h, w->pp, st);
    for (int pr = 0; pr < 4; pr++)
      for (int cc = 0; cc < 2; cc++)
        PNFAM_CUDA_CHECK(cudaMemcpyAsync(w->stage.p + (size_t)(8 + 2 * pr + cc) * nxy, w->hsp.p + ((size_t)cc * 4 + quad_of_pair[pr]) * nxy,
                                         (size_t)out[2 * pr + cc].nelem * sizeof(double), cudaMemcpyDeviceToHost, st));
    PNFAM_CUDA_CHECK(cudaStreamSynchronize(st));
    PNFAM_CUDA_CHECK(cudaGetLastError());
    for (int i = 0; i < 8; i++) std::memcpy(out[i].elem, w->stage.p + (size_t)(8 + i) * nxy, (size_t)out[i].nelem * sizeof(double));
    c->launches += 3 + 1 + 5;
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

// ---- DMMA peak probe ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; j++) { c[j][0] = 0.0; c[j][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) dmma884(c[j][0], c[j][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- self-check of the fused transform plan (host only, no device needed) ----------------------------------
// Builds the operator plan for the given block structure, regroups its tasks into fused jobs and evaluates both forms
// with random operands on the CPU.  out[0..3] = forward {tasks, jobs, products(jobs)/products(tasks), max |difference|},
// out[4..7] = the same for the backward transform.  Returns 0, or 2 if the structure is left to the two-phase kernels.
extern "C" int pnfam_b200_check_transform_plan(int32_t nb, const int32_t* db_, const int32_t* f_ir2c, int32_t use_diag, int32_t beta_minus,
                                               double* out, char* err, int errlen) {
  try {
    std::vector<int> db(db_, db_ + nb), ir2c(f_ir2c, f_ir2c + nb);
    OperatorPlan op = make_operator_plan(db, ir2c, use_diag != 0, beta_minus != 0);
    size_t wsize = 0;
    for (int d : db) wsize += (size_t)d * d;
    unsigned long long seed = 88172645463325252ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed >> 11) / 9007199254740992.0 - 0.5; };
    std::vector<double> W[4];
    for (auto& w : W) { w.resize(wsize); for (double& v : w) v = rnd(); }
    const size_t nxy = op.nxy;
    std::vector<double> in((size_t)2 * 4 * nxy);
    for (double& v : in) v = rnd();
    int rc = 0;
    for (int dir = 0; dir < 2; dir++) {
      const TransformPlan& tp = dir == 0 ? op.forward : op.backward;
      std::vector<FusedJob> jobs;
      std::vector<int4> ctas[3];
      if (!build_fused_host(tp, jobs, ctas)) { rc = 2; out[4 * dir] = (double)tp.tasks.size(); out[4 * dir + 1] = 0; continue; }
      auto elemA = [&](int mat, int off, int tr, int m, int i, int k) { return tr ? W[mat][off + k + (size_t)i * m] : W[mat][off + i + (size_t)k * m]; };
      auto elemB = [&](int c, int quad, int off, int tr, int m, int n, int k, int j) {
        const double* B = in.data() + ((size_t)c * 4 + quad) * nxy + off;
        return tr ? B[j + (size_t)k * n] : B[k + (size_t)j * m];
      };
      auto elemC = [&](int mat, int off, int tr, int n, int k, int j) { return tr ? W[mat][off + j + (size_t)k * n] : W[mat][off + k + (size_t)j * n]; };
      std::vector<double> ref((size_t)2 * 4 * nxy, 0.0), got((size_t)2 * 4 * nxy, 0.0);
      double p_old = 0, p_new = 0;
      for (const BlockTask& b : tp.tasks) {
        const int m = b.m, n = b.n;
        p_old += (double)b.nterms * ((double)m * m * n + (double)m * n * n);
        for (int c = 0; c < 2; c++)
          for (int t = 0; t < b.nterms; t++) {
            const TripleTerm& x = b.t[t];
            std::vector<double> T((size_t)m * n, 0.0);
            for (int i = 0; i < m; i++)
              for (int j = 0; j < n; j++) {
                double sacc = 0;
                for (int k = 0; k < m; k++) sacc += elemA(x.a_mat, x.a_off, x.a_trans, m, i, k) * elemB(c, x.b_quad, x.b_off, x.b_trans, m, n, k, j);
                T[i + (size_t)j * m] = sacc;
              }
            const double al = c ? x.alpha_im : x.alpha_re;
            for (int i = 0; i < m; i++)
              for (int j = 0; j < n; j++) {
                double sacc = 0;
                for (int k = 0; k < n; k++) sacc += T[i + (size_t)k * m] * elemC(x.c_mat, x.c_off, x.c_trans, n, k, j);
                ref[((size_t)c * 4 + b.out_quad) * nxy + b.out_off + i + (size_t)j * m] += al * sacc;
              }
          }
      }
      for (const FusedJob& J : jobs) {
        const int m = J.m, n = J.n;
        for (int x = 0; x < J.ngroups; x++) p_new += (double)J.nterms[x] * m * m * n + (double)J.nout * m * n * n;
        for (int c = 0; c < 2; c++)
          for (int x = 0; x < J.ngroups; x++) {
            std::vector<double> T((size_t)m * n, 0.0);
            for (int t = 0; t < J.nterms[x]; t++) {
              const FusedTerm& f = J.t[x][t];
              const double al = c ? f.alpha_im : f.alpha_re;
              for (int i = 0; i < m; i++)
                for (int j = 0; j < n; j++) {
                  double sacc = 0;
                  for (int k = 0; k < m; k++) sacc += elemA(f.a_mat, f.a_off, f.a_trans, m, i, k) * elemB(c, f.b_quad, f.b_off, f.b_trans, m, n, k, j);
                  T[i + (size_t)j * m] += al * sacc;
                }
            }
            for (int o = 0; o < J.nout; o++) {
              const FusedOut& fo = J.o[o];
              const double be = c ? fo.beta_im[x] : fo.beta_re[x];
              for (int i = 0; i < m; i++)
                for (int j = 0; j < n; j++) {
                  double sacc = 0;
                  for (int k = 0; k < n; k++) sacc += T[i + (size_t)k * m] * elemC(fo.c_mat[x], fo.c_off[x], fo.c_trans[x], n, k, j);
                  got[((size_t)c * 4 + fo.out_quad) * nxy + fo.out_off + i + (size_t)j * m] += be * sacc;
                }
            }
          }
      }
      double worst = 0;
      for (size_t i = 0; i < ref.size(); i++) worst = std::max(worst, std::fabs(ref[i] - got[i]));
      out[4 * dir] = (double)tp.tasks.size(); out[4 * dir + 1] = (double)jobs.size(); out[4 * dir + 2] = p_new / p_old; out[4 * dir + 3] = worst;
    }
    return rc;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

extern "C" int pnfam_b200_dmma_peak(int device, double* tflops, char* err, int errlen) {
  try {
    DeviceGuard dg(device);
    cudaDeviceProp prop;
    PNFAM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    DBuf<double> out;
    const int blocks = prop.multiProcessorCount * 4, iters = 20000;
    out.alloc((size_t)blocks * 256);
    cudaEvent_t e0, e1;
    PNFAM_CUDA_CHECK(cudaEventCreate(&e0)); PNFAM_CUDA_CHECK(cudaEventCreate(&e1));
    dmma_peak_kernel<<<blocks, 256>>>(out.p, 1000);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
      PNFAM_CUDA_CHECK(cudaEventRecord(e0));
      dmma_peak_kernel<<<blocks, 256>>>(out.p, iters);
      PNFAM_CUDA_CHECK(cudaEventRecord(e1));
      PNFAM_CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = (double)blocks * 8 /*warps*/ * iters * 8 * 512.0;
      best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}
