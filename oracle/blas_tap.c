/* Test infrastructure -- not part of the product.
 *
 * BLAS/LAPACK "tap": a stand-in libblas.so.3 / liblapack.so.3 for the prebuilt reference executables
 * (oracle/_ref/pnfam_main.x) that forwards every call to the real OpenBLAS and, when PNFAM_TAP=<file>
 * is set, appends the arguments of the dgemm_ and dsyevr_ calls to <file>.  The reference binary has
 * no debug dumps; its dgemm operands ARE the intermediates we need to pin (wave-function tables at
 * pnfam_hamiltonian_blas.f90:201-236, U/V/X blocks at pnfam_type_blockmatrix.f90:202-206, the
 * hpsi arrays at pnfam_hamiltonian_blas.f90:1132-1167; dsyevr operands are the HFB blocks of
 * hfbtho_solver.f90:1575).  Records are read back by oracle/tapfile.py.
 *
 * Record layout (little-endian):  int32 kind (1=dgemm, 2=dsyevr), then
 *  dgemm : int32 ta,tb,m,n,k,lda,ldb,ldc ; f64 alpha,beta ; A (ra*ca f64, packed col-major),
 *          B (rb*cb), C_after (m*n)
 *  dsyevr: int32 n, lda, m_found, ldz, uplo ; A_before (n*n) ; w (n) ; Z (n*m_found)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *real_lib;
static FILE *tapf;
static int tap_init_done;
static long tap_max_bytes = 1L << 31, tap_bytes;

static void tap_init(void) {
  if (tap_init_done) return;
  tap_init_done = 1;
  real_lib = RTLD_NEXT; /* the real OpenBLAS is a NEEDED entry of this library (see Makefile) */
  const char *f = getenv("PNFAM_TAP");
  if (f && *f) tapf = fopen(f, "wb");
  const char *mx = getenv("PNFAM_TAP_MAX");
  if (mx) tap_max_bytes = atol(mx);
}
static void *sym(const char *n) {
  tap_init();
  void *p = dlsym(real_lib, n);
  if (!p) { fprintf(stderr, "blas_tap: missing %s\n", n); exit(2); }
  return p;
}

typedef long L;
#define ARGS L a1, L a2, L a3, L a4, L a5, L a6, L a7, L a8, L a9, L a10, L a11, L a12, L a13, L a14, \
             L a15, L a16, L a17, L a18, L a19, L a20, L a21, L a22, L a23, L a24, L a25, L a26, L a27, L a28
#define PASS a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, \
             a21, a22, a23, a24, a25, a26, a27, a28
/* All Fortran BLAS/LAPACK arguments are pointers (+ by-value hidden string lengths): forwarding 28
 * integer-class words is ABI-safe on x86-64 SysV (extra stack words are ignored by the callee). */
#define FWD_VOID(name) \
  void name(ARGS) { static void (*f)(ARGS); if (!f) f = (void (*)(ARGS))sym(#name); f(PASS); }
#define FWD_DBL(name) \
  double name(ARGS) { static double (*f)(ARGS); if (!f) f = (double (*)(ARGS))sym(#name); return f(PASS); }

FWD_VOID(daxpy_) FWD_VOID(dcopy_) FWD_VOID(dscal_) FWD_VOID(dgemv_) FWD_VOID(zaxpy_) FWD_VOID(zgemm_)
FWD_VOID(dgetrf_) FWD_VOID(dgetri_) FWD_VOID(zgetrf_) FWD_VOID(zgetri_) FWD_VOID(dsytrf_) FWD_VOID(dsytri_)
FWD_DBL(ddot_) FWD_DBL(dnrm2_) FWD_DBL(dlamch_)

static void put_i(int v) { fwrite(&v, 4, 1, tapf); tap_bytes += 4; }
static void put_d(double v) { fwrite(&v, 8, 1, tapf); tap_bytes += 8; }
static void put_mat(const double *a, int rows, int cols, int ld) {
  for (int j = 0; j < cols; j++) fwrite(a + (size_t)j * ld, 8, rows, tapf);
  tap_bytes += 8L * rows * cols;
}

void dgemm_(const char *ta, const char *tb, const int *m, const int *n, const int *k, const double *alpha,
            const double *a, const int *lda, const double *b, const int *ldb, const double *beta, double *c,
            const int *ldc, size_t l1, size_t l2) {
  static void (*f)(const char *, const char *, const int *, const int *, const int *, const double *,
                   const double *, const int *, const double *, const int *, const double *, double *,
                   const int *, size_t, size_t);
  if (!f) f = sym("dgemm_");
  f(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, l1, l2);
  if (tapf && tap_bytes < tap_max_bytes) {
    int tA = (*ta == 'T' || *ta == 't'), tB = (*tb == 'T' || *tb == 't');
    put_i(1); put_i(tA); put_i(tB); put_i(*m); put_i(*n); put_i(*k); put_i(*lda); put_i(*ldb); put_i(*ldc);
    put_d(*alpha); put_d(*beta);
    put_mat(a, tA ? *k : *m, tA ? *m : *k, *lda);
    put_mat(b, tB ? *n : *k, tB ? *k : *n, *ldb);
    put_mat(c, *m, *n, *ldc);
    fflush(tapf);
  }
}

void dsyevr_(const char *jobz, const char *range, const char *uplo, const int *n, double *a, const int *lda,
             const double *vl, const double *vu, const int *il, const int *iu, const double *abstol, int *m,
             double *w, double *z, const int *ldz, int *isuppz, double *work, const int *lwork, int *iwork,
             const int *liwork, int *info, size_t l1, size_t l2, size_t l3) {
  static void (*f)(const char *, const char *, const char *, const int *, double *, const int *, const double *,
                   const double *, const int *, const int *, const double *, int *, double *, double *,
                   const int *, int *, double *, const int *, int *, const int *, int *, size_t, size_t, size_t);
  if (!f) f = sym("dsyevr_");
  int dump = tapf && tap_bytes < tap_max_bytes && *lwork != -1 && *liwork != -1;
  double *acopy = NULL;
  if (dump) {
    acopy = malloc(sizeof(double) * (size_t)(*n) * (*n));
    for (int j = 0; j < *n; j++) memcpy(acopy + (size_t)j * (*n), a + (size_t)j * (*lda), 8 * (size_t)(*n));
  }
  f(jobz, range, uplo, n, a, lda, vl, vu, il, iu, abstol, m, w, z, ldz, isuppz, work, lwork, iwork, liwork, info,
    l1, l2, l3);
  if (dump) {
    put_i(2); put_i(*n); put_i(*lda); put_i(*m); put_i(*ldz); put_i((int)*uplo);
    put_mat(acopy, *n, *n, *n); fwrite(w, 8, *n, tapf); tap_bytes += 8L * (*n);
    put_mat(z, *n, *m, *ldz);
    fflush(tapf);
    free(acopy);
  }
}
