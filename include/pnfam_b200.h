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
 *              init_external_field            exes/pnfam/pnfam_extfield.f90:37-108.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnfam_problem pnfam_problem;

/* rundir holds hfbtho_NAMELIST.dat + hfbtho_output.hel (+ <name>.tbc); namelist_file is the pnFAM
 * namelist (relative to rundir unless absolute) -- the argv[1] of pnfam_main.x
 * (exes/pnfam/pnfam_setup.f90:214-250). */
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

#ifdef __cplusplus
}
#endif
#endif /* PNFAM_B200_H */
