// Host-callable launchers of the FAM iteration kernels.
#pragma once
#include "device_common.cuh"

namespace pnfam {

// ---- batch control (device resident) --------------------------------------------------------------
// A solve keeps S "slots" (work space of one omega point each) busy: when the point of a slot converges, the next
// pending point is admitted into it by batch_control_kernel (mixer.cu) without the host taking part.  Every kernel of
// the iteration indexes its work by za < nactive through active[za] = slot; the host sizes the grids with a (stale)
// upper bound of nactive and CTAs beyond the device-side count return at once.
struct BatchCtrl {
  int nactive;        // slots running an iteration
  int next_pending;   // position in the admission order of the next point to admit
  int ndone;          // points finished (converged or interrupted at max_iter)
  int step;           // lock-step iterations executed so far
};

// a per-device high-water mark (shared-memory attributes are per device: a process may hold contexts on several GPUs)
struct PerDeviceMax {
  size_t v[64] = {};
  bool raise(size_t x) {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (x <= v[d]) return false;
    v[d] = x;
    return true;
  }
};

// ---- (a)/(d) transforms ---------------------------------------------------------------------------
struct DevicePlan {
  const DevTask* tasks = nullptr;
  const int4* tiles1 = nullptr;   // phase 1: (task, term, first row, first column) of every non-empty 32x32 tile
  const int4* tiles2 = nullptr;   // phase 2: (task, 0, first row, first column)
  int ntasks = 0, ntiles1 = 0, ntiles2 = 0, max_dim = 0;
  size_t scratch_elems = 0;       // doubles of scratch per (point, re/im)
  // fused path (one launch per size class, no scratch): jobs + (job, first m-tile, m-tiles, 0) per CTA, largest first
  const FusedJob* jobs = nullptr;
  const int4* fctas[3] = {nullptr, nullptr, nullptr};   // size classes: n <= 88, <= 128, <= 176
  int nfctas[3] = {0, 0, 0};
  bool fused = false;
  int njobs = 0;
};

struct TransformArgs {
  const double* W[4];             // Ua, Va, Ub, Vb block arrays
  const double* in;               // middle operand, per point
  size_t in_pstride;
  int in_pack;                    // 1: Broyden pack layout, 0: [c][quad][nxy]
  double* out;
  size_t out_pstride;
  int out_pack;
  double* scratch;                // [nactive][2][scratch_stride]
  size_t scratch_stride;
  size_t nxy;
  const int* active;              // [nactive] slot indices
  const BatchCtrl* ctrl;          // device-side active count (nullptr: the grid is exact)
};

// returns the number of kernels launched
int launch_transform(const DevicePlan& plan, const TransformArgs& args, int nactive, cudaStream_t stream);

// ---- (b)/(c) hamiltonian --------------------------------------------------------------------------
// Bilinear grid densities D[t][t'][s][s'] = sum_{a in s, b in s'} phi^t_a(r) rho_ab phi^t'_b(r)
constexpr int NDD_RHO = 4 * 4 * 2 * 2 * 2;  // doubles per grid point (complex)
constexpr int NDD_KAP = 2 * 2 * 2;
constexpr int NMF = 5 * 5 * 2 * 2 * 2;       // field tensor mf(ta,tb,sa,sb) complex, doubles per grid point
constexpr int NPF = 2 * 2 * 2;               // pairing field (sa,sb) complex
// Both field tensors are stored tile-major so that the projection fetches what one CTA needs for one r-tile with a
// single linear bulk copy:
//   mf[kt][sa][sb][pair][rr][c]     (the 18 structurally non-zero (ta,tb) pairs: 1152 doubles per (r-tile kt,
//                                    row spin sa); a column chunk inside one spin segment fetches one sb half)
//   pf[sa][kt][sb][rr][c]           (64 doubles per (sa, kt); kt padded to a multiple of 4: one copy per 4 r-tiles)
constexpr int MF_PAIRS = 18;
constexpr int MF_TILE = 2 * MF_PAIRS * RT * 2;
constexpr int PF_TILE = 2 * RT * 2;
__host__ __device__ inline size_t mf_elems(int ntiles) { return (size_t)ntiles * 2 * MF_TILE; }
__host__ __device__ inline size_t pf_elems(int ntiles) { return (size_t)2 * ((ntiles + 3) & ~3) * PF_TILE; }

// One pipeline step of the density kernel: an (a-chunk x b-chunk) piece of one block of rho / kappa.
// Rows / columns of a block are spin-sorted (up first).  A chunk may straddle the spin boundary when the padded
// segments fit (small blocks): the step then runs up to four (s, s') sub-passes on one staged operand set.
struct DensStep {
  int a_row0, na_up, na_dn;   // first row of the contraction chunk in the padded index space; valid rows per spin
  int b_row0, nb_up, nb_dn;   // first row of the column chunk in the padded index space; valid columns per spin
  int rho_off, ld;            // element offset of (a chunk start, b chunk start) inside the block matrix, leading dim
  int flags;                  // bit0: this step brings in its phi_a image (else it shares the one of an earlier
                              // step); bit1: first a-chunk (zero C); bit2: last a-chunk (epilogue)
  int kp;                     // row stride of the packed chunk: >= padded a-count, kp % 8 == 4 (bank-conflict free)
  int pk_off;                 // offset (doubles) of the packed chunk [n = (b,c)][k = a] in the packed rho array
  // shared-memory ring schedule, simulated on the host (build_density_steps): phi_a lives at aoff, phi_b and the rho
  // chunk at [soff, ..) of the arena; they may be written once step `dep` has been released by all warps; when the
  // math reaches this step, steps [.., issue_to) are due.
  int soff, aoff;             // byte offsets in the arena
  int dep;                    // last step whose arena space / barrier slot this step reuses (-1: none)
  int issue_to;               // steps [.., issue_to) are issued when the math reaches this step
  int pad;
};
constexpr int DENS_ARENA = 220 * 1024;   // bytes of shared memory cycled through by the density kernel
constexpr int DENS_NBAR = 16;            // mbarrier slots (step k uses slot k % DENS_NBAR)
constexpr int DENS_LOOKAHEAD = 2;        // steps issued ahead of the math (measured: deeper queues of bulk copies delay
                                         // the operands that are needed next; 2 is the optimum on B200)
constexpr int DENS_MAXSTEPS = 2048;      // steps of one density pass (their dependency list lives in shared memory)
constexpr int DENS_AC = 48;   // contraction chunk
constexpr int DENS_BC = 32;   // column chunk

// ---- sum-factorised path (hamiltonian_sf.cu): used when the model carries the separable factors of the basis ----
// phi^t_a(ih, il) = Z(zrow_a, ih) R_a(il):  the contraction over a runs over the radial index first (FMA, small) and
// over n_z second (DMMA with K = number of distinct n_z of the spin segment), see hamiltonian_sf.cu.
constexpr int SF_KMAX = 16;      // distinct n_z ("slots") per spin segment
constexpr int SF_SEGTAB = 40;    // ints per segment table: [0..16] first row of slot k (rows sorted by slot; [nslots] = n),
                                 // [17..32] z-table row of slot k, [33] nslots, [34] n
constexpr int SF_THREADS = 512;
constexpr int SF_MFP = 18;       // structurally non-zero (t,t') pairs of the field tensor
struct SfDensStep {              // one (a spin segment) x (b column chunk inside one spin segment) piece of a block
  int seg_a, a_row0, na, nslots; // segment index 2*block+s, first padded row, states, distinct n_z
  int b_row0, nbc;               // first padded row of the column chunk, padded columns (multiple of 4)
  int img_off;                   // offset (doubles) of the packed rho image [na][2 nbc] of this step
  int sweep;                     // s*2 + s'  (the kernel accumulates one (s,s') combination at a time)
  int flags;                     // bit0: last step of its sweep
  int rho_off, ld;               // element offset of the block in the block matrix, leading dimension
  int pad;
};
struct SfProjTile {              // pair task of the projection: (a spin segment) x (4 padded columns b of one spin segment);
                                 // na = 0: padding entry (the list holds groups of 8 tasks with equal (sa, sb))
  int seg_a, a_row0, na, nslots;
  int b_row0, nbc, sa, sb;
  int out_off, ld, pad0, pad1;
};
struct SfDev {
  int enabled = 0;
  int ngh = 0, ngl = 0, mt = 0;  // mt: DMMA m-tiles (8 grid points) per il
  int kih = 0;                   // ngh padded to a multiple of 4
  int zs = 0;                    // row stride of the z tables (doubles), zs % 16 == 4
  int nzrows = 0, dqp_p = 0;
  const double* zt = nullptr;    // [3][nzrows][zs]  Z0, Z1, Z2 (zero beyond ngh)
  const double* rg = nullptr;    // [ngl][dqp_p][4]  R0..R3 per padded row (rows of a segment sorted by n_z slot)
  const double* rgp = nullptr;   // [(ngl+1)/2][dqp_p][2][4]  the same for il pairs (density: one copy serves both il)
  const double* r0q = nullptr;   // [(ngl+3)/4][dqp_p][4]  R0 of four consecutive il (pairing density: one pass over kappa serves four il)
  const double* rgt = nullptr;   // [ngl][4][dqp_p]  component-major copy of rg (radial projection: consecutive lanes = consecutive
                                 // rows read consecutive doubles)
  const int* zrow = nullptr;     // [dqp_p] z-table row of a padded row (0 for padding)
  const int* p2l = nullptr;      // [dqp_p] index of the state inside its block, -1 for padding
  const int* slot = nullptr;     // [dqp_p] n_z slot of the row inside its spin segment
  const int* segtab = nullptr;   // [2 nb][SF_SEGTAB]
  const SfDensStep* steps[4] = {nullptr, nullptr, nullptr, nullptr};   // rho q0, rho q1, kappa q0, kappa q1
  int nsteps[4] = {0, 0, 0, 0};
  size_t pk_stride[2] = {0, 0};  // doubles of packed images per (point, pass): rho, kappa
  double* pk[2] = {nullptr, nullptr};
  int na_max = 0, nbc_max = 0, kpad_max = 0;
  const SfProjTile* tiles[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [h / Delta][pass]
  int ntiles[2][2] = {{0, 0}, {0, 0}};
  int ksplit = 1;
};
// ---- fully factorised path (hamiltonian_sf2.cu) ----------------------------------------------------------------
// phi^t_a(ih, il) = Z^m(n_z(a), ih) R^j_a(il) on BOTH sides of every grid contraction:
//   density     D^{tt'}_{ss'}(ih,il) = sum_{zr,zr'} Z^m(zr,ih) Z^m'(zr',ih) Pi^{jj'}_{ss'}[zr][zr'][il],
//               Pi^{jj'}[zr][zr'][il] = sum_{a in (s,zr), b in (s',zr')} R^j_a(il) rho_ab R^j'_b(il)       (radial part, O(nxy ngl))
//   projection  kt^{jj'}_{sa sb}[zr][zr'][il] = sum_{(t,t')->(j,j')} sum_ih Z^m(zr,ih) mf^{tt'}(ih,il) Z^m'(zr',ih)
//               h_ab = 2 sum_il sum_{jj'} R^j_a(il) kt^{jj'}[zr_a][zr_b][il] R^j'_b(il)                      (radial part, O(nxy ngl))
// (zr: row of the z table = distinct n_z of the whole basis, <= N_sh + 3).  The number of executed flops no longer
// scales with ngh * nxy: ~10x fewer than the one-sided factorisation of hamiltonian_sf.cu at 16 shells.
constexpr int SF2_NJJ = 11;      // (j, j') combinations of the radial factors in the mean field:
                                 // (0,0) (0,1) (0,2) (0,3) (1,0) (1,1) (1,2) (2,0) (2,1) (2,2) (3,0)
constexpr int SF2_RUN = 8;       // columns per task of the radial projection (runs of equal n_z are cut at this length)
struct Sf2Task {                 // radial projection: one or two rows a of equal n_z x a run of <= SF2_RUN columns with equal n_z
  int pa, pb0, nb;               // padded row of (the first) a, of the first column, columns
  int out_base, ld;              // element offset of the block in the block matrix, leading dimension
  int sasb;                      // spin combination 2 sa + sb
  int na;                        // rows: 1 or 2 (pa, pa + 1: the kt entries of an il are fetched once for both)
  int pad;
};
struct Sf2Dev {
  int enabled = 0;
  int nzr = 0;
  // density: the elements of rho / kappa regrouped by (sweep, zr, zr') -- element i of list k (rho q0, rho q1, kappa q0,
  // kappa q1) is rsp[..][el_src[k][i]]; the packed copy pk[..][i] = (re, im) is written in this order every iteration.
  // cols[k][c] = (first element, rows, padded row of the first row, padded row of the column): the columns of the
  // sub-blocks of a group (contiguous runs of the packed copy); cptr: first column of every (sweep, zr, zr') group.
  const int* el_src[4] = {nullptr, nullptr, nullptr, nullptr};
  const int4* cols[4] = {nullptr, nullptr, nullptr, nullptr};
  const int* cptr[4] = {nullptr, nullptr, nullptr, nullptr};          // [4 sweeps][nzr*nzr + 1]
  int nelem[4] = {0, 0, 0, 0};
  const int2* zrange[4] = {nullptr, nullptr, nullptr, nullptr};       // [4 sweeps][nzr]: zr' range with non-empty pair lists
  const int* order[4] = {nullptr, nullptr, nullptr, nullptr};          // [4 sweeps][1 + nzr*nzr]: count, then the non-empty (zr, zr')
                                                                       // entries by decreasing work (balanced dealing to the lanes)
  const Sf2Task* tasks[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [h / Delta][pass]
  int ntasks[2][2] = {{0, 0}, {0, 0}};
  const unsigned char* need[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [h / Delta][pass]: [4 sasb][nzr*nzr] kt entries in use
  double* kt[2] = {nullptr, nullptr};  // [slot za][2 q][4 sasb][ngl][nzr*nzr][NJJ or 1][2]
};
__host__ __device__ inline size_t sf2_kt_elems(int mode, int ngl, int nzr) { return (size_t)4 * ngl * nzr * nzr * (mode == 0 ? SF2_NJJ : 1) * 2; }

__host__ __device__ inline size_t sf_mf_elems(int ngl, int kih) { return (size_t)ngl * 4 * SF_MFP * kih * 2; }
constexpr int SF_DIL = 4;       // Gauss-Laguerre nodes per iteration of the Delta projection (their work is a quarter of h's)
// pf[sa][sb][il][ih][c] with SF_DIL rows of zero padding per (sa, sb): SF_DIL consecutive nodes are one linear copy
__host__ __device__ inline size_t sf_pf_elems(int ngl, int kih) { return (size_t)4 * (ngl + SF_DIL) * kih * 2; }

struct SideStreams;
struct HamArgs {
  DevBasis basis;
  SfDev sf;
  Sf2Dev sf2;
  // per pass q=0 (pn,+) / q=1 (np,-): input structures (dRsp quadrants) and output structures (dHsp quadrants)
  DevBlockStruct rho_in[2], kap_in[2], h_out[2], d_out[2];
  int rho_quad[2], kap_quad[2];   // storage quadrant of rho / kappa (and of h / Delta) for each pass
  const DensStep* steps_rho[2]; int nsteps_rho[2];   // pipeline step lists of the density kernel, per pass
  const DensStep* steps_kap[2]; int nsteps_kap[2];
  double* pk_rho;                 // [nactive][2 q][pk_stride_rho]  rho chunks repacked step by step (pack_rho_kernel)
  double* pk_kap;                 // [nactive][2 q][pk_stride_kap]
  size_t pk_stride_rho, pk_stride_kap;
  const double* rsp;              // [P][2 c][4][nxy]
  double* hsp;                    // [P][2 c][4][nxy]
  size_t nxy;
  double* dd_rho;                 // [nactive][2 q][NDD_RHO][nghl]
  double* dd_kap;                 // [nactive][2 q][NDD_KAP][nghl]
  double* mf;                     // [nactive][2 q][mf_elems]  tile-major, see above
  double* pf;                     // [nactive][2 q][pf_elems]
  double* hpart;                  // split-K partials of the projection
  const int* active;
  int nactive;                    // grid size (upper bound of the device-side count when ctrl is given)
  const BatchCtrl* ctrl;          // nullptr: nactive is exact
  SideStreams* side;              // side streams of the owning context (host side only)
};

struct ProjPlan {                 // output tiles of the grid->HO projection
  const int4* tiles_h = nullptr;  // (block row, a-chunk start (padded index space), b-chunk start, spin bits)
  const int4* tiles_d = nullptr;
  int ntiles_h[2] = {0, 0}, ntiles_d[2] = {0, 0};
  int tile_off_h[2] = {0, 0}, tile_off_d[2] = {0, 0};
  int ksplit = 1;
};

// Independent kernels of one stage (rho and kappa densities; h and Delta projections of both passes) run on side
// streams forked from / joined to the caller's stream, so that the tail of one fills with CTAs of the next.  Every
// context owns its set (streams and events belong to the device that was current at creation; two contexts never share
// fork / join events).
struct SideStreams {
  static constexpr int N = 3;
  cudaStream_t s[N];
  cudaEvent_t fork, join[N];
  SideStreams() {
    for (int i = 0; i < N; i++) {
      PNFAM_CUDA_CHECK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
      PNFAM_CUDA_CHECK(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
    }
    PNFAM_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  }
  ~SideStreams() {
    for (int i = 0; i < N; i++) { cudaStreamDestroy(s[i]); cudaEventDestroy(join[i]); }
    cudaEventDestroy(fork);
  }
  SideStreams(const SideStreams&) = delete;
  SideStreams& operator=(const SideStreams&) = delete;
  void fork_from(cudaStream_t main, int n) {
    PNFAM_CUDA_CHECK(cudaEventRecord(fork, main));
    for (int i = 0; i < n; i++) PNFAM_CUDA_CHECK(cudaStreamWaitEvent(s[i], fork, 0));
  }
  void join_to(cudaStream_t main, int n) {
    for (int i = 0; i < n; i++) {
      PNFAM_CUDA_CHECK(cudaEventRecord(join[i], s[i]));
      PNFAM_CUDA_CHECK(cudaStreamWaitEvent(main, join[i], 0));
    }
  }
};

void launch_density(const HamArgs& a, cudaStream_t stream);
// host helper: flatten a block structure into density pipeline steps
void build_density_steps(int nb, const int* db, const int* pstart, const int* nsu, const int* r2c, const int* r2m,
                         DensStep* out, int* nout, size_t* pk_elems);  // out may be null to count
void launch_fields(const HamArgs& a, cudaStream_t stream);
void launch_projection(const HamArgs& a, const ProjPlan& pp, cudaStream_t stream);
size_t projection_partial_elems(const ProjPlan& pp, size_t nxy);
// sum-factorised variants (hamiltonian_sf.cu)
void launch_density_sf(const HamArgs& a, cudaStream_t stream);
void launch_projection_sf(const HamArgs& a, cudaStream_t stream);
void launch_density_sf2(const HamArgs& a, cudaStream_t stream);     // hamiltonian_sf2.cu
void launch_projection_sf2(const HamArgs& a, cudaStream_t stream);
void launch_sf_pack(const HamArgs& a, cudaStream_t stream);          // packed rho / kappa images (hamiltonian_sf.cu)
int sf2_smem_bytes(const SfDev& S);
int sf_density_smem_bytes(const SfDev& S);      // dynamic shared memory the kernels need for this basis
int sf_projection_smem_bytes(const SfDev& S);

// ---- (d) Greens function update, Broyden mixer, strength -----------------------------------------
struct MixArgs {
  int nvec;                       // 4 (X,Y re/im) or 8 (+P,Q)
  size_t nxy, n;                  // n = nvec*nxy
  int M;                          // allocated history slots (stride of df/dv/gram), >= 1
  int Mmode;                      // the namelist's broyden_history_size: <0 no mixing, 0 linear, >0 Broyden
  double alpha;                   // mixing factor
  double w0;
  const double* hqp;              // [P][2][4][nxy]
  const double* fqp;              // [4][nxy]
  const double* esum;             // [4][nxy]  b*f1_i + c*f2_j of matrix_2qp for each qp quadrant
  const double* tfac;             // [4][nxy] or nullptr
  const double* omega;            // [P][2]
  double quench;
  double* vin;                    // [P][n]
  double* vout;                   // [P][n]
  double* df;                     // [P][M][n]
  double* dv;                     // [P][M][n]
  double* du;                     // [P][M][n]  dv + alpha df of the normalised history entries: the update streams ONE array
  double* gram;                   // [P][M][M]
  double* work;                   // [P][M]   df_i . vout
  double* gamma;                  // [P][M]
  double* dotpart;                // [P][M][nslices][2] slice partials of (df_i . df_new, df_i . vout)
  int nslices;                    // broyden_slices(n)
  double* red;                    // [P][nred][2] partial (max, sumsq)
  int nred;
  double* si;                     // [P]
  double* normi;                  // [P]
  const double* gqp;              // [1+nx][4][nxy] (F first, then cross-terms)
  int nstr;                       // 1 + nxterms
  double* strength;               // [P][nstr][2]
  double* strpart;                // [nactive][nstr][STR_SPLIT][2] slices of the strength sums
  double* chol;                   // [S][M][M+1] Cholesky work space in global memory, used when the history outgrows shared memory
  const int* active;
  int nactive;                    // grid size (upper bound of the device-side count)
  const BatchCtrl* ctrl;
  const int* slot_iter;           // [S] iterations the point of a slot has completed (the `iter` of qrpa_broyden)
};
void launch_reset(const MixArgs& a, cudaStream_t stream);     // zero the amplitudes of freshly admitted slots
void launch_greens(const MixArgs& a, cudaStream_t stream);
int broyden_slices(size_t n);
size_t strength_partial_elems(int npoints, int nstr);
void launch_broyden(const MixArgs& a, cudaStream_t stream);
void launch_strength(const MixArgs& a, cudaStream_t stream);
int broyden_launches(const MixArgs& a);

// Retire / admit (one CTA, end of every lock-step iteration): for every active slot, count the iteration, publish
// (iters, si, strengths, trace row) of its point, decide convergence (si < eps) or interruption (max_iter), hand free
// slots to pending points in admission order, compact the active list.
struct BatchArgs {
  BatchCtrl* ctrl;
  int* active;                    // [S]
  int* slot_iter;                 // [S]
  int* slot_point;                // [S] point held by a slot
  const int* order;               // [P] admission order (point indices)
  int npoints, nslots, max_iter, nstr;
  double eps;
  const double* si;               // [S]
  const double* strength;         // [S][nstr][2]
  double* omega;                  // [S][2]   frequency of the point a slot holds
  const double* omega_pt;         // [P][2]
  // per-point results
  int* out_iters;                 // [P]
  int* out_conv;                  // [P]
  double* out_si;                 // [P]
  double* out_strength;           // [P][nstr][2]
  double* out_trace;              // [P][max_iter+1][4] or nullptr: si, Re S, Im S, lock-step index of the iteration
};
void launch_batch_control(const BatchArgs& b, cudaStream_t stream);

}  // namespace pnfam
