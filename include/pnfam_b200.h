/* pnfam_b200.h -- C ABI of the B200-native pnFAM response solver.
 *
 * Two shared libraries implement it (built by `make -C pynfam_b200/csrc`):
 *   libpnfam_host.so  (CPU only)  : section 1, the host set-up a Fortran caller already owns
 *   libpnfam_b200.so  (CUDA sm_100a): section 2, the FAM iteration -- THE HOT PATH
 *
 * All pointers are plain host pointers unless stated; all matrices are Fortran column-major,
 * indices in *_ir2c / *_ir2m are 1-based exactly as in the reference's `type(blockmatrix)`
 * (exes/pnfam/pnfam_type_blockmatrix.f90:15-26), so a Fortran `bind(C)` interface can pass its own
 * arrays without copies (INTEGRATION.md shows the ISO_C_BINDING block).
 * Every entry point returns 0 on success; on failure a message is copied to err (if given).
 */
#ifndef PNFAM_B200_H
#define PNFAM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * 1. Host set-up (stand-in for the reference's Fortran set-up; one-off per nucleus / operator)
 *    replaces: setup_pnfam + setup_extfield   exes/pnfam/pnfam_solver.f90:44-48,
 *              get_raw_hfb_solution           exes/pnfam/hfbtho_solution.f90:137-395,
 *              HFBTHO_program (0 iterations)  exes/pnfam/hfbtho_interface.f90:20-226,
 *              init_interaction               exes/pnfam/pnfam_interaction.f90:90-260,
 *              init_external_field            exes/pnfam/pnfam_extfield.f90:37-108,
 *              effective_2bc_extfield         exes/pnfam/pnfam_extfield_2bc.f90:26-465 (+ write_tbc / read_tbc,
 *                                             exes/pnfam/pnfam_storage.f90:488-727).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnfam_problem pnfam_problem;

/* rundir holds hfbtho_NAMELIST.dat + hfbtho_output.hel (+ <name>.tbc); namelist_file is the pnFAM
 * namelist (relative to rundir unless absolute) -- the argv[1] of pnfam_main.x
 * (exes/pnfam/pnfam_setup.f90:214-250).  Full-FAM two-body currents (two_body_current_mode = x1x1xx): <name>.tbc is read
 * when it fits the calculation, otherwise the field is computed and the file (re)written in the reference's record
 * layout, as the reference does (PNFAM_B200_NO_TBC_GENERATOR=1: fail instead). */
int pnfam_problem_create(const char* rundir, const char* namelist_file, pnfam_problem** out, char* err, int errlen);
/* Same, but reuses the HFB reconstruction of an existing problem (same rundir / nucleus). */
int pnfam_problem_create_shared(const pnfam_problem* nucleus_of, const char* rundir, const char* namelist_file,
                                pnfam_problem** out, char* err, int errlen);
void pnfam_problem_destroy(pnfam_problem* p);
/* Named read-only views (valid until destroy).  Names: see pynfam_b200/host.py. */
int pnfam_problem_scalar(const pnfam_problem* p, const char* name, double* out);
int pnfam_problem_array_f64(pnfam_problem* p, const char* name, const double** ptr, int64_t* n);
int pnfam_problem_array_i32(pnfam_problem* p, const char* name, const int32_t** ptr, int64_t* n);
/* which: 0 operator label, 1..nxterms cross-term labels, -1 interaction name, -2 output base name */
int pnfam_problem_label(const pnfam_problem* p, int which, char* out, int outlen);
/* The Yukawa part of the full-FAM two-body-current Gamow-Teller field from the arrays the reference's own routine works on:
 * replaces  effective_2bc_extfield(op)  exes/pnfam/pnfam_extfield_2bc.f90:26-465, called by setup_extfield
 * (exes/pnfam/pnfam_solver.f90:622-640) when <name>.tbc cannot be read.  Inputs are what that routine takes from the modules
 * hfb_solution / hfbtho / type_blockmatrix (exes/pnfam/hfbtho_solution.f90:40-58):
 *   ntx, nbx        states and blocks of the HFBTHO basis (Omega > 0 only);  id[nbx] = states per block
 *   hnz, hnr, hnl, hns [ntx]   n_z, n_r, Lambda, 2 s of every state (hfbtho's nz, nr, nl, ns)
 *   bz, bp          oscillator lengths
 *   rmat            HO-basis density matrix rk: column-major (ld_rmat, 2 nbx); column ib holds the id(ib)^2 elements of the
 *                   neutron block ib, column nbx + ib of the proton block (rho_db = rmat / 2)
 *   ir2c, ir2m [2 nbx]   1-based block structure of op%mat in the doubled basis (type_blockmatrix);  nxy = size(op%mat%elem)
 *   k, beta_minus   operator K (-1, 0, +1) and beta type;  use_p = two_body_current_usep
 *   spin_sorted     1: rows / columns of every block in the spin-sorted order of the USE_HBLAS = 1 build, 0: original order
 * Outputs (each [nxy], may be NULL): the six parts of op%gam with the low-energy constants stripped, exactly the records
 * write_tbc stores (exes/pnfam/pnfam_storage.f90:526-534):  op%mat%elem = c3 (c3d + c3e) + (c4 + 1/4) (c4d + c4e) + cpd + cpe. */
int pnfam_host_effective_2bc_extfield(int32_t ntx, int32_t nbx, const int32_t* id, const int32_t* hnz, const int32_t* hnr,
                                      const int32_t* hnl, const int32_t* hns, double bz, double bp, const double* rmat,
                                      int64_t ld_rmat, const int32_t* ir2c, const int32_t* ir2m, int64_t nxy, int32_t k,
                                      int32_t beta_minus, int32_t use_p, int32_t spin_sorted, double* c3d, double* c3e,
                                      double* c4d, double* c4e, double* cpd, double* cpe, char* err, int errlen);
/* OpenMP threads of the host set-up (n <= 0: unchanged); returns the previous setting.  The reconstructed HFB solution
 * is kept in <rundir>/.pnfam_b200_hfb_<hash of the two input files>.cache (PNFAM_B200_CACHE_DIR: another directory,
 * PNFAM_B200_NO_CACHE: off): later launches in directories holding the same two files load it instead of repeating
 * gamdel / hfbdiag / DENSIT. */
int pnfam_host_set_threads(int n);

/* ------------------------------------------------------------------------------------------------
 * 2. The FAM iteration on the GPU (libpnfam_b200.so, CUDA sm_100a) -- THE HOT PATH
 *    replaces: ifam                    exes/pnfam/pnfam_solver.f90:93-221   (whole loop, batched over omega)
 *              init_pnfam_solver       exes/pnfam/pnfam_solver.f90:226-460  (F -> qp basis, Greens, T)
 *              calc_hamiltonian (ptr)  exes/pnfam/pnfam_setup.f90:56-71, pnfam_hamiltonian_blas.f90:51-74
 *              triprod_bbm             exes/pnfam/pnfam_type_bbm.f90:557-576
 *              qrpa_broyden            exes/pnfam/pnfam_broyden.f90:96-216
 *    There is NO CPU fallback: every entry point fails with an error if no CUDA device is usable.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnfam_b200_ctx pnfam_b200_ctx;

/* Nucleus + residual interaction: the module variables of hfb_solution / pnfam_interaction /
 * type_blockmatrix that the reference's iteration reads.  Tables are (nghl, dqp) column-major with the
 * rows of every block sorted spin-up first (hfbtho_solution.f90:309-351). */
typedef struct {
  int32_t nb, dqp, nghl;
  const int32_t* db;            /* [nb]  block dimensions                                   */
  const int32_t* num_spin_up;   /* [nb]                                                     */
  const double *wf, *wfdr, *wfdp, *wfdz, *wfd2_all;  /* [nghl*dqp]                          */
  const double *wdcori, *crho, *cs, *cpair, *cspair; /* [nghl]                              */
  double cdrho, ctau, ctj0, ctj1, ctj2, crdj, cds, ct, cj, cgs, cf, csdj;
  const double *Ep, *En;        /* [dqp]  quasiparticle energies (0 above the pairing window) */
  const double *Up, *Vp, *Un, *Vn; /* [sum db^2] block-diagonal storage (pnfam_setup.f90:292-321) */
  const double *qp_fp, *qp_fn;  /* [dqp] equal-filling / thermal occupations, or NULL        */
  /* Optional separable description of the same tables (harmonic-oscillator basis; HFBTHO builds them as
   * products QH(nz,ih) * QL(nr,Lambda,il), hfbtho_solver.f90:3463-3671, hfbtho_solution.f90:199-202, 346-351):
   *   ihil = ih + il*ngh (0-based),  z = sep_zrow[state]
   *   wf = Z0[z][ih] R0[state][il]   wfdr = Z0 R1   wfdp = Z0 R2   wfdz = Z1 R0   wfd2_all = Z2 R0 + Z0 R3
   * When given (ngh > 0) the density / projection kernels contract the two grid directions one after the other
   * (sum factorisation) and the full tables are only used to verify the factors.  ngh = 0: general tables. */
  int32_t ngh, ngl, sep_nzrows;
  const int32_t* sep_zrow;      /* [dqp]                                                      */
  const double* sep_z;          /* [3][sep_nzrows][ngh]   Z0, Z1, Z2                          */
  const double* sep_r;          /* [4][dqp][ngl]          R0, R1, R2, R3                      */
} pnfam_b200_model;

int pnfam_b200_ctx_create(const pnfam_b200_model* model, int device, pnfam_b200_ctx** out, char* err, int errlen);
void pnfam_b200_ctx_destroy(pnfam_b200_ctx* ctx);
/* 1 if the context runs the sum-factorised kernels (separable factors given and accepted), 0 if the general-table ones */
int pnfam_b200_ctx_separable(const pnfam_b200_ctx* ctx);
/* Bytes ctx_create copied host -> device (basis tables or their separable factors, U, V, grid couplings). */
int64_t pnfam_b200_ctx_h2d_bytes(const pnfam_b200_ctx* ctx);

/* External field f (+ cross-term fields g_k sharing f's block structure), single-particle basis,
 * as produced by init_external_field / setup_crossterms (pnfam_extfield.f90:37-108, 882-949). */
typedef struct {
  int32_t beta_minus;           /* 1: beta-, 0: beta+                                        */
  int32_t nxterms;
  const int32_t* f_ir2c;        /* [nb] 1-based partner column block, 0 = none               */
  const double* f_elem;         /* [nxy]                                                     */
  const double* const* g_elem;  /* [nxterms][nxy]                                            */
} pnfam_b200_operator;

typedef struct {
  int32_t max_iter;             /* &solver max_iter                                          */
  int32_t broyden_history_size; /* &solver broyden_history_size (any size, as the reference)  */
  double convergence_epsilon, quench_residual_int, energy_shift_prot, energy_shift_neut;
  int32_t batch_slots;          /* omega points iterated side by side (work space = slots x per-point state);
                                   0 = automatic: min(npoints, 64, what fits the free device memory).  Points beyond
                                   the slot count are admitted as running ones finish; results do not depend on it */
  int32_t reserved;             /* 0                                                         */
} pnfam_b200_solver_params;

typedef struct {
  double seconds_total;         /* wall time of the call                                     */
  double seconds_device;        /* CUDA-event time of the iteration loop                     */
  int64_t iterations;           /* sum over points of FAM iterations performed               */
  int64_t kernel_launches;      /* kernels launched by this call                             */
  int64_t h2d_bytes, d2h_bytes;
  double seconds_density, seconds_projection; /* CUDA-event time spent in the two dominant kernels */
  int64_t launches_density, launches_projection;
  double flops_density, flops_projection;     /* algorithmic FP64 flops of those launches: (16+4) and (20+4) x nghl x nxy
                                                 per (point, pass), SURVEY.md section 8d */
  int32_t batch_slots;          /* slots the solve ran with                                   */
  int32_t lock_steps;           /* lock-step iterations of the batch (>= max over points of iters) */
} pnfam_b200_stats;

/* Solve npoints complex frequencies of one operator.  Outputs (host buffers):
 *   strength [npoints][1+nxterms][2]  S and cross-terms (re,im), as the "Result" table of the .dat
 *   iters    [npoints]                iteration count at exit (iter_conv of ifam), or max_iter
 *   conv     [npoints]                1 converged, 0 interrupted
 *   si       [npoints]                last max|dX|
 *   trace    [npoints][max_iter+1][4] optional (may be NULL): si, Re S, Im S, seconds per iteration */
int pnfam_b200_solve(pnfam_b200_ctx* ctx, const pnfam_b200_operator* op, const pnfam_b200_solver_params* prm,
                     int32_t npoints, const double* omega_re, const double* omega_im, double* strength,
                     int32_t* iters, int32_t* conv, double* si, double* trace, pnfam_b200_stats* stats, char* err,
                     int errlen);

/* The reference's plug-in point `calc_hamiltonian` (pnfam_setup.f90:56-71): 8 input and 8 output
 * block matrices in the reference's argument order
 *   in : rerho_pn imrho_pn rekp imkp rerho_np imrho_np rekm imkm
 *   out: reh_pn imh_pn redp imdp reh_np imh_np redm imdm
 * The caller presets ir2c/ir2m of the outputs, exactly as pnfam_solver.f90:402-413 does. */
typedef struct {
  double* elem;                 /* [nelem]                                                   */
  const int32_t* ir2c;          /* [nb] 1-based, 0 = none                                    */
  const int32_t* ir2m;          /* [nb] 1-based offsets                                      */
  int64_t nelem;
} pnfam_b200_blockmatrix;
int pnfam_b200_calc_hamiltonian(pnfam_b200_ctx* ctx, const pnfam_b200_blockmatrix in[8], pnfam_b200_blockmatrix out[8],
                                char* err, int errlen);

/* FP64 DMMA (mma.sync m8n8k4) peak probe used for the roofline denominator of the tensor-bound
 * kernels: returns achieved TFLOP/s of a register-resident DMMA loop on all SMs. */
int pnfam_b200_dmma_peak(int device, double* tflops, char* err, int errlen);

/* Host-only self-check of the transform plan (no device needed): builds the block-task list of
 * triprod_bbm (pnfam_type_bbm.f90:428-576) for the given block structure, regroups it into the jobs of the
 * fused transform kernel and evaluates both forms on the CPU with random operands.
 * out[0..3] = forward {tasks, jobs, products(jobs) / products(tasks), max |difference|}, out[4..7] = backward.
 * Returns 0, 2 if the structure stays with the two-phase kernels, 1 on error. */
int pnfam_b200_check_transform_plan(int32_t nb, const int32_t* db, const int32_t* f_ir2c, int32_t use_diag,
                                    int32_t beta_minus, double* out, char* err, int errlen);

#ifdef __cplusplus
}
#endif
#endif /* PNFAM_B200_H */
