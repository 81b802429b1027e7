/* Test infrastructure (oracle/_ref runtime shim) -- not part of the product.
 * The prebuilt reference executables import exactly one GFORTRAN_10 symbol,
 * _gfortran_os_error_at, which the GFORTRAN_8-era libgfortran bundled in this image lacks.
 * Everything else resolves from that real libgfortran (pulled in through OpenBLAS' NEEDED). */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
void _gfortran_os_error_at(const char *where, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "%s: ", where ? where : "?");
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
  exit(1);
}
void _shim_dummy8(void) {}
