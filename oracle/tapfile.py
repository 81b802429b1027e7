"""Test infrastructure -- reader for the records written by oracle/blas_tap.c."""
import numpy as np


def read_tap(path, max_records=None):
    recs = []
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            kind = int(np.frombuffer(h, "<i4")[0])
            if kind == 1:
                ta, tb, m, n, k, lda, ldb, ldc = np.frombuffer(f.read(32), "<i4")
                alpha, beta = np.frombuffer(f.read(16), "<f8")
                ra, ca = (k, m) if ta else (m, k)
                rb, cb = (n, k) if tb else (k, n)
                A = np.frombuffer(f.read(8 * ra * ca), "<f8").reshape(ca, ra).T
                B = np.frombuffer(f.read(8 * rb * cb), "<f8").reshape(cb, rb).T
                C = np.frombuffer(f.read(8 * m * n), "<f8").reshape(n, m).T
                recs.append(dict(kind="dgemm", ta=int(ta), tb=int(tb), m=int(m), n=int(n), k=int(k),
                                 alpha=float(alpha), beta=float(beta), A=A, B=B, C=C))
            elif kind == 2:
                n, lda, m, ldz, uplo = np.frombuffer(f.read(20), "<i4")
                A = np.frombuffer(f.read(8 * n * n), "<f8").reshape(n, n).T
                w = np.frombuffer(f.read(8 * n), "<f8")
                Z = np.frombuffer(f.read(8 * n * m), "<f8").reshape(m, n).T
                recs.append(dict(kind="dsyevr", n=int(n), m=int(m), uplo=chr(uplo), A=A, w=w, Z=Z))
            else:
                raise ValueError("bad tap record kind %d" % kind)
            if max_records and len(recs) >= max_records:
                break
    return recs
