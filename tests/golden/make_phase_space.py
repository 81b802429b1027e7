#!/usr/bin/env python
"""Golden values of the reference's phase-space module (pynfam/strength/phase_space.py) for the branches the beta.out
fixtures do not reach: beta-plus (negative Z, Rose screening), the F_1 Fermi function, lambda_2, real-axis integrals and the
Thiele continuation at complex W0.  The reference module is imported from /root/reference with the numpy-1 aliases it
still uses; run ONCE in the build container, tests/golden/phase_space.json is committed."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

np.float_, np.complex_ = np.float64, np.complex128
for name, path in (("pynfam", "/root/reference/pynfam"), ("pynfam.strength", "/root/reference/pynfam/strength")):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
for name, path in (("pynfam.config", "/root/reference/pynfam/config.py"), ("pynfam.strength.phase_space", "/root/reference/pynfam/strength/phase_space.py")):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
ps = sys.modules["pynfam.strength.phase_space"]

w = np.array([1.0, 1.02, 1.5, 3.0, 8.0, 20.0])
rec = {"w": w.tolist(), "fermi": [], "lambda2": [], "calcPsi": [], "psiFct": []}
f = lambda a: [repr(float(x)) for x in np.atleast_1d(a)]
for Z, A, sc in ((17, 40, False), (65, 162, False), (-15, 40, True), (-63, 162, True), (-15, 40, False)):
    for F in (0, 1):
        rec["fermi"].append({"F": F, "Z": Z, "A": A, "sc": sc, "val": f(ps.Fermi(F, Z, A, w.copy(), sc))})
    rec["lambda2"].append({"Z": Z, "A": A, "sc": sc, "val": f(ps.lambda_ke(2, Z, A, w.copy(), sc))})
w0 = np.array([1.0, 1.3, 4.0, 12.0, 21.5])
zc = np.array([2.0 + 0.5j, 10.0 - 3.0j, 18.0 + 0.1j])
for beta, Z, A in (("-", 17, 40), ("+", 15, 40), ("+", 63, 162), ("-", 65, 162)):
    p = ps.phaseSpace(beta)
    Zd = Z if beta == "-" else -Z
    sc = beta == "+"
    for n in range(1, 7):
        rec["calcPsi"].append({"beta": beta, "Z": Zd, "A": A, "sc": sc, "n": n, "w0": w0.tolist(), "val": f(p.calcPsi(n, Zd, A, w0, sc=sc))})
        fct = p.psiFct(n, Zd, A, 10.0, 0.0, approx=True)
        v = fct(zc)
        rec["psiFct"].append({"beta": beta, "Z": Zd, "A": A, "n": n, "eqrpamax": 10.0, "eqrpamin": 0.0,
                              "re_z": np.real(zc).tolist(), "im_z": np.imag(zc).tolist(), "re": f(np.real(v)), "im": f(np.imag(v))})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "phase_space.json")
json.dump({"source": "mld1812/pynfam pynfam/strength/phase_space.py, imported in the build container", **rec}, open(dst, "w"), indent=0)
print(dst, os.path.getsize(dst))
