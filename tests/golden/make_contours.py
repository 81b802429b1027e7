#!/usr/bin/env python
"""Golden contour geometries from the reference's own famContour (pynfam/strength/contour.py), all seven contour types
with default and overridden settings -> tests/golden/contours.json.  The reference module is imported from
/root/reference with the numpy-1 aliases it still uses; run ONCE in the build container, the JSON is committed."""
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
for name, path in (("pynfam.config", "/root/reference/pynfam/config.py"), ("pynfam.strength.contour", "/root/reference/pynfam/strength/contour.py")):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
famContour = sys.modules["pynfam.strength.contour"].famContour

CASES = [
    ("CIRCLE", None), ("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036}),
    ("CIRCLE", {"energy_min": -1.0, "energy_max": 80.0, "nr_points": 21, "use_gauleg_ctr": False}),
    ("CIRCLE", {"energy_max": 9.0, "nr_points": 15, "shift_imag": 0.2}),
    ("CONSTL", None), ("CONSTL", {"energy_min": 1.0, "energy_max": 4.0, "nr_points": 7, "half_width": 0.25}),
    ("CONSTR", None), ("CONSTR", {"energy_max": 3.0, "half_width": 0.2, "de_hw_ratio": 0.5}),
    ("FERMIS", None), ("FERMIS", {"energy_min": 0.5, "energy_max": 9.0, "hw_min": 0.02, "hw_max": 0.5, "u_percent_interval": 0.4}),
    ("EXP", None), ("EXP", {"energy_max": 12.0, "p_percent_interval": 0.3, "de_hw_ratio": 1.5}),
    ("MONOMIAL", None), ("MONOMIAL", {"energy_max": 7.0, "power": 2.0, "hw_max": 0.3}),
    ("FERMIA", None), ("FERMIA", {"energy_max": 10.0, "nr_points_max": 40}), ("FERMIA", {"energy_max": 30.0, "nr_points_max": 60}),
    ("FERMIA", {"energy_max": 2.0, "nr_points_max": 300}), ("FERMIA", {"energy_max": 40.0, "nr_points_max": 50}),
]
out = []
for name, ov in CASES:
    c = famContour(name, ov)
    out.append({"name": name, "override": ov, "nr_points": int(c.nr_points), "nr_compute": int(c.nr_compute), "closed": bool(c.closed),
                "quadrature": c.quadrature, "name_and_int": c.name_and_int,
                "re": [repr(float(x)) for x in np.real(c.ctr_z)], "im": [repr(float(x)) for x in np.imag(c.ctr_z)],
                "dzdt_re": [repr(float(x)) for x in np.real(c.ctr_dzdt)], "dzdt_im": [repr(float(x)) for x in np.imag(c.ctr_dzdt)]})
    print(name, ov, c.nr_points)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "contours.json")
json.dump({"source": "mld1812/pynfam pynfam/strength/contour.py famContour, imported in the build container", "cases": out}, open(dst, "w"), indent=0)
print(dst, os.path.getsize(dst))
