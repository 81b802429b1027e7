"""ctypes view of libpnfam_b200.so (section 2 of include/pnfam_b200.h): the FAM iteration on the GPU.

There is no CPU fallback: if the CUDA library is missing or no device is usable these calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class Model(ctypes.Structure):
    _fields_ = ([("nb", ctypes.c_int32), ("dqp", ctypes.c_int32), ("nghl", ctypes.c_int32),
                 ("db", c_int32_p), ("num_spin_up", c_int32_p)]
                + [(k, c_double_p) for k in ("wf", "wfdr", "wfdp", "wfdz", "wfd2_all",
                                             "wdcori", "crho", "cs", "cpair", "cspair")]
                + [(k, ctypes.c_double) for k in ("cdrho", "ctau", "ctj0", "ctj1", "ctj2", "crdj",
                                                  "cds", "ct", "cj", "cgs", "cf", "csdj")]
                + [(k, c_double_p) for k in ("Ep", "En", "Up", "Vp", "Un", "Vn", "qp_fp", "qp_fn")]
                + [("ngh", ctypes.c_int32), ("ngl", ctypes.c_int32), ("sep_nzrows", ctypes.c_int32),
                   ("sep_zrow", c_int32_p), ("sep_z", c_double_p), ("sep_r", c_double_p)])


class Operator(ctypes.Structure):
    _fields_ = [("beta_minus", ctypes.c_int32), ("nxterms", ctypes.c_int32), ("f_ir2c", c_int32_p),
                ("f_elem", c_double_p), ("g_elem", ctypes.POINTER(c_double_p))]


class SolverParams(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int32), ("broyden_history_size", ctypes.c_int32),
                ("convergence_epsilon", ctypes.c_double), ("quench_residual_int", ctypes.c_double),
                ("energy_shift_prot", ctypes.c_double), ("energy_shift_neut", ctypes.c_double),
                ("batch_slots", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [("seconds_total", ctypes.c_double), ("seconds_device", ctypes.c_double),
                ("iterations", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
                ("seconds_density", ctypes.c_double), ("seconds_projection", ctypes.c_double),
                ("launches_density", ctypes.c_int64), ("launches_projection", ctypes.c_int64),
                ("flops_density", ctypes.c_double), ("flops_projection", ctypes.c_double),
                ("batch_slots", ctypes.c_int32), ("lock_steps", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BlockMatrixC(ctypes.Structure):
    _fields_ = [("elem", c_double_p), ("ir2c", c_int32_p), ("ir2m", c_int32_p), ("nelem", ctypes.c_int64)]


class GpuError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "libpnfam_b200.so")
        if not os.path.isfile(path):
            raise GpuError("libpnfam_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(path)
        vp, cp, ci = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int
        L.pnfam_b200_ctx_create.argtypes = [ctypes.POINTER(Model), ci, ctypes.POINTER(vp), cp, ci]
        L.pnfam_b200_ctx_destroy.argtypes = [vp]
        L.pnfam_b200_ctx_destroy.restype = None
        L.pnfam_b200_ctx_separable.argtypes = [vp]
        L.pnfam_b200_ctx_h2d_bytes.argtypes = [vp]
        L.pnfam_b200_ctx_h2d_bytes.restype = ctypes.c_int64
        L.pnfam_b200_solve.argtypes = [vp, ctypes.POINTER(Operator), ctypes.POINTER(SolverParams), ctypes.c_int32,
                                       c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p, c_double_p,
                                       c_double_p, ctypes.POINTER(Stats), cp, ci]
        L.pnfam_b200_calc_hamiltonian.argtypes = [vp, ctypes.POINTER(BlockMatrixC), ctypes.POINTER(BlockMatrixC), cp, ci]
        L.pnfam_b200_dmma_peak.argtypes = [ci, c_double_p, cp, ci]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


class Context:
    """Device-resident nucleus + interaction (pnfam_b200_ctx)."""

    def __init__(self, problem, device=0, separable=True):
        """separable=False withholds the separable factors of the basis tables: the general-table kernels run."""
        L = lib()
        p = problem
        self._keep = k = {"problem": problem}    # the model points into the problem's own arrays (no host copies)
        m = Model()
        m.nb, m.dqp, m.nghl = p.iscalar("nb"), p.iscalar("dqp"), p.iscalar("nghl")
        k["db"] = np.ascontiguousarray(p.i32("db"))
        k["nsu"] = np.ascontiguousarray(p.i32("num_spin_up"))
        m.db, m.num_spin_up = _ip(k["db"]), _ip(k["nsu"])
        for name in ("wf", "wfdr", "wfdp", "wfdz", "wfd2_all", "wdcori", "crho", "cs", "cpair", "cspair",
                     "Ep", "En", "Up", "Vp", "Un", "Vn"):
            k[name] = p.f64(name, copy=False)
            setattr(m, name, _dp(k[name]))
        for name in ("cdrho", "ctau", "ctj0", "ctj1", "ctj2", "crdj", "cds", "ct", "cj", "cgs", "cf", "csdj"):
            setattr(m, name, p.scalar(name))
        if p.iscalar("statistical"):      # odd-A equal filling or finite temperature: P,Q quadrants, T factors
            k["qp_fp"], k["qp_fn"] = np.ascontiguousarray(p.f64("qp_fp")), np.ascontiguousarray(p.f64("qp_fn"))
            m.qp_fp, m.qp_fn = _dp(k["qp_fp"]), _dp(k["qp_fn"])
        if separable:
            m.ngh, m.ngl, m.sep_nzrows = p.iscalar("ngh"), p.iscalar("ngl"), p.iscalar("sep_nzrows")
            k["sep_zrow"] = np.ascontiguousarray(p.i32("sep_zrow"))
            k["sep_z"], k["sep_r"] = p.f64("sep_z", copy=False), p.f64("sep_r", copy=False)
            m.sep_zrow, m.sep_z, m.sep_r = _ip(k["sep_zrow"]), _dp(k["sep_z"]), _dp(k["sep_r"])
        self.nb = m.nb
        self._h = ctypes.c_void_p()
        err = ctypes.create_string_buffer(1024)
        if L.pnfam_b200_ctx_create(ctypes.byref(m), device, ctypes.byref(self._h), err, 1024) != 0:
            self._h = None
            raise GpuError(err.value.decode())

    @property
    def separable(self):
        """True when the sum-factorised kernels run (the model carried usable separable factors)."""
        return bool(lib().pnfam_b200_ctx_separable(self._h))

    @property
    def h2d_bytes(self):
        """Bytes the context creation copied host -> device."""
        return int(lib().pnfam_b200_ctx_h2d_bytes(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pnfam_b200_ctx_destroy(self._h)
            self._h = None

    def solve(self, problem, omegas=None, max_iter=None, eps=None, history=None, want_trace=False, slots=0):
        """Solve the problem's operator at the given complex frequencies (default: the namelist's one).
        slots: omega points iterated side by side (0 = automatic); the others are admitted as running points finish.
        Returns dict(strength[P, 1+nx] complex, iters, conv, si, stats, labels, trace)."""
        L = lib()
        p = problem
        if omegas is None:
            omegas = [complex(p.scalar("real_eqrpa"), p.scalar("imag_eqrpa"))]
        om = np.asarray(omegas, dtype=complex)
        P = len(om)
        wre, wim = np.ascontiguousarray(om.real), np.ascontiguousarray(om.imag)
        nx = p.iscalar("nxterms")
        f_ir2c, f_elem = np.ascontiguousarray(p.i32("f_ir2c")), np.ascontiguousarray(p.f64("f_elem"))
        g = [np.ascontiguousarray(p.f64("g_elem_%d" % i)) for i in range(nx)]
        for i in range(nx):
            if not np.array_equal(p.i32("g_ir2c_%d" % i), f_ir2c):
                raise GpuError("cross-term field has a different block structure than the operator")
        garr = (c_double_p * max(nx, 1))(*[_dp(x) for x in g])
        op = Operator(int(p.iscalar("beta_minus")), nx, _ip(f_ir2c), _dp(f_elem), garr)
        prm = SolverParams(int(max_iter if max_iter is not None else p.iscalar("max_iter")),
                           int(history if history is not None else p.iscalar("broyden_history_size")),
                           float(eps if eps is not None else p.scalar("convergence_epsilon")),
                           p.scalar("quench_residual_int"), p.scalar("energy_shift_prot"), p.scalar("energy_shift_neut"),
                           int(slots), 0)
        strength = np.zeros((P, 1 + nx, 2))
        iters, conv = np.zeros(P, np.int32), np.zeros(P, np.int32)
        si = np.zeros(P)
        trace = np.zeros((P, prm.max_iter + 1, 4)) if want_trace else None
        stats = Stats()
        err = ctypes.create_string_buffer(1024)
        rc = L.pnfam_b200_solve(self._h, ctypes.byref(op), ctypes.byref(prm), P, _dp(wre), _dp(wim), _dp(strength),
                                _ip(iters), _ip(conv), _dp(si), _dp(trace) if want_trace else None,
                                ctypes.byref(stats), err, 1024)
        if rc != 0:
            raise GpuError(err.value.decode())
        return dict(strength=strength[..., 0] + 1j * strength[..., 1], iters=iters, conv=conv, si=si,
                    stats=stats.as_dict(), trace=trace,
                    labels=[p.label(i) for i in range(1 + nx)])

    def calc_hamiltonian(self, ins, outs):
        """ins/outs: 8 oracle-style block matrices each (objects with .elem, .ir2c, .ir2m; 1-based maps).
        Fills outs[i].elem in place (the reference's calc_hamiltonian interface)."""
        keep = []

        def conv(b):
            e = np.ascontiguousarray(b.elem, dtype=np.float64)
            c, m_ = np.ascontiguousarray(b.ir2c, dtype=np.int32), np.ascontiguousarray(b.ir2m, dtype=np.int32)
            keep.extend([e, c, m_])
            return BlockMatrixC(_dp(e), _ip(c), _ip(m_), len(e)), e
        cin = (BlockMatrixC * 8)()
        cout = (BlockMatrixC * 8)()
        outbuf = []
        for i in range(8):
            cin[i], _ = conv(ins[i])
            cout[i], e = conv(outs[i])
            outbuf.append(e)
        err = ctypes.create_string_buffer(1024)
        if lib().pnfam_b200_calc_hamiltonian(self._h, cin, cout, err, 1024) != 0:
            raise GpuError(err.value.decode())
        for i in range(8):
            outs[i].elem = outbuf[i]


def dmma_peak_tflops(device=0):
    v = ctypes.c_double()
    err = ctypes.create_string_buffer(1024)
    if lib().pnfam_b200_dmma_peak(device, ctypes.byref(v), err, 1024) != 0:
        raise GpuError(err.value.decode())
    return v.value
