"""ctypes view of libpnfam_host.so (section 1 of include/pnfam_b200.h): the host set-up that the
reference performs in Fortran before entering the FAM iteration (pnfam_solver.f90:44-62)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "libpnfam_host.so")
        if not os.path.isfile(path):
            raise RuntimeError("libpnfam_host.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
                               " or `make -C pynfam_b200/csrc`")
        L = ctypes.CDLL(path)
        vp, cp, ci = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int
        L.pnfam_problem_create.argtypes = [cp, cp, ctypes.POINTER(vp), cp, ci]
        L.pnfam_problem_create_shared.argtypes = [vp, cp, cp, ctypes.POINTER(vp), cp, ci]
        L.pnfam_problem_destroy.argtypes = [vp]
        L.pnfam_problem_destroy.restype = None
        L.pnfam_problem_scalar.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_double)]
        L.pnfam_problem_array_f64.argtypes = [vp, cp, ctypes.POINTER(ctypes.POINTER(ctypes.c_double)),
                                              ctypes.POINTER(ctypes.c_int64)]
        L.pnfam_problem_array_i32.argtypes = [vp, cp, ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                              ctypes.POINTER(ctypes.c_int64)]
        L.pnfam_problem_label.argtypes = [vp, ci, cp, ci]
        L.pnfam_host_set_threads.argtypes = [ci]
        i32, i64, dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_double
        pi32, pdbl = ctypes.POINTER(i32), ctypes.POINTER(dbl)
        L.pnfam_host_effective_2bc_extfield.argtypes = [i32, i32, pi32, pi32, pi32, pi32, pi32, dbl, dbl, pdbl, i64, pi32, pi32, i64,
                                                        i32, i32, i32, i32, pdbl, pdbl, pdbl, pdbl, pdbl, pdbl, cp, ci]
        _LIB = L
    return _LIB


class PnfamError(RuntimeError):
    pass


def set_threads(n):
    """OpenMP threads of the host set-up (n <= 0: unchanged); returns the previous setting."""
    return int(lib().pnfam_host_set_threads(int(n)))


def effective_2bc_extfield(id_, nz, nr, nl, ns, bz, bp, rmat, ir2c, ir2m, nxy, k, beta_minus=True, use_p=False, spin_sorted=True):
    """The reference's effective_2bc_extfield (pnfam_extfield_2bc.f90:26-465) on plain arrays, through the C ABI
    (include/pnfam_b200.h: pnfam_host_effective_2bc_extfield).  rmat: HFBTHO's rk as a Fortran-ordered (nqx, 2 nbx) array.
    Returns the six LEC-stripped parts (c3d, c3e, c4d, c4e, cpd, cpe) as an array of shape (6, nxy)."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    id_, nz, nr, nl, ns, ir2c, ir2m = map(i32, (id_, nz, nr, nl, ns, ir2c, ir2m))
    rmat = np.asfortranarray(rmat, dtype=np.float64)
    out = np.zeros((6, int(nxy)))
    err = ctypes.create_string_buffer(1024)
    pi = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    pd = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rc = lib().pnfam_host_effective_2bc_extfield(len(nz), len(id_), pi(id_), pi(nz), pi(nr), pi(nl), pi(ns), float(bz), float(bp),
                                                 pd(rmat), rmat.shape[0], pi(ir2c), pi(ir2m), int(nxy), int(k), int(bool(beta_minus)),
                                                 int(bool(use_p)), int(bool(spin_sorted)), pd(out[0]), pd(out[1]), pd(out[2]),
                                                 pd(out[3]), pd(out[4]), pd(out[5]), err, 1024)
    if rc != 0:
        raise PnfamError(err.value.decode(errors="replace"))
    return out


class Problem:
    """One (nucleus, operator, K, omega) problem: what `pnfam_main.x <namelist>` sets up in `rundir`."""

    def __init__(self, rundir, namelist, share_nucleus_with=None):
        L = lib()
        self._h = ctypes.c_void_p()
        err = ctypes.create_string_buffer(1024)
        if share_nucleus_with is None:
            rc = L.pnfam_problem_create(os.fsencode(rundir), os.fsencode(namelist), ctypes.byref(self._h), err, 1024)
        else:
            rc = L.pnfam_problem_create_shared(share_nucleus_with._h, os.fsencode(rundir), os.fsencode(namelist),
                                               ctypes.byref(self._h), err, 1024)
        if rc != 0:
            self._h = None
            raise PnfamError(err.value.decode())
        self.rundir = rundir
        self.namelist = namelist

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pnfam_problem_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def scalar(self, name):
        v = ctypes.c_double()
        if lib().pnfam_problem_scalar(self._h, name.encode(), ctypes.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    def iscalar(self, name):
        return int(round(self.scalar(name)))

    def f64(self, name, copy=True):
        """Named array; copy=False returns a read-only view of the library's own storage (valid while this Problem
        is alive) -- what gpu.Context passes straight to the C ABI, as a Fortran caller would pass its module arrays."""
        p = ctypes.POINTER(ctypes.c_double)()
        n = ctypes.c_int64()
        if lib().pnfam_problem_array_f64(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n)) != 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0)
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        if copy:
            return a.copy()
        a.flags.writeable = False
        return a

    def i32(self, name):
        p = ctypes.POINTER(ctypes.c_int32)()
        n = ctypes.c_int64()
        if lib().pnfam_problem_array_i32(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n)) != 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def label(self, which=0):
        buf = ctypes.create_string_buffer(256)
        if lib().pnfam_problem_label(self._h, which, buf, 256) != 0:
            raise KeyError(which)
        return buf.value.decode()

    def table(self, name):
        """(nghl, dqp) wave-function table as a Fortran-ordered 2-D array."""
        return self.f64(name).reshape(self.iscalar("dqp"), self.iscalar("nghl")).T
