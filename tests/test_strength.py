"""Contour geometry and strength assembly (pynfam_b200/strength.py) against the reference's own OP.out / OP.out.ctr
files (tests/golden/*/fam_soln, copied from tests/pynfam_test_S40 and tests/S40_GT_All by tests/golden/make_strength.py).
CPU only: the file formats are byte-exact, the contour to round-off of the Gauss-Legendre nodes."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN
from pynfam_b200.strength import beta_rate, complex_quadrature, famContour, famStrength, patch_namelist

SOLN = {"GT-": os.path.join(GOLDEN, "S40_SKOP_6sh", "fam_soln"),     # tests/pynfam_test_S40
        "RS1-": os.path.join(GOLDEN, "S40_GT_All", "fam_soln"),      # tests/S40_GT_All (current cross-term definitions)
        "PS0-": os.path.join(GOLDEN, "S40_GT_All", "fam_soln")}
EMAX = 10.476036   # EQRPA_max of the S-40 HFB solution ("CIRCLE on (0.00e+00, 1.05e+01)" in the fixture headers)
OPS = [("GT-", 0), ("GT-", 1), ("RS1-", 1), ("PS0-", 0)]


def _fixture(op, k):
    fs = famStrength(op, k, "CIRCLE")
    fs.readCtrBinary(SOLN[op])
    return fs


@pytest.mark.parametrize("op,k", OPS)
def test_circle_contour_matches_the_reference_files(op, k):
    ref = _fixture(op, k).contour
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": EMAX})
    assert c.nr_points == 60 and c.nr_compute == 30 and c.closed and c.use_gauleg
    for name in ("theta", "ctr_z", "ctr_dzdt", "glwts"):
        a, b = getattr(c, name), getattr(ref, name)
        assert np.max(np.abs(a - b)) < 1e-12 * max(1.0, np.max(np.abs(b))), name
    assert abs(np.sum(c.glwts) - 2 * np.pi) < 1e-12


@pytest.mark.parametrize("op,k", OPS)
def test_ctr_binary_round_trip_is_byte_exact(op, k, tmp_path):
    fs = _fixture(op, k)
    assert fs.nucleus == (24, 16, 40) and fs.version == 3
    fs.writeCtrBinary(str(tmp_path))
    name = fs.file_bin
    assert open(os.path.join(str(tmp_path), name), "rb").read() == open(os.path.join(SOLN[op], name), "rb").read()


@pytest.mark.parametrize("op,k", OPS)
def test_strength_out_text_is_byte_exact(op, k, tmp_path):
    """Feed the reference's numbers through concatFamData (computed half only) + writeStrengthOut."""
    fs = _fixture(op, k)
    lines = open(os.path.join(SOLN[op], fs.file_txt)).read().split("\n")
    rows = [ln.split() for ln in lines[9:] if ln.strip()][:30]
    full = fs.cstr_df.values
    fresh = famStrength(op, k, famContour("CIRCLE", {"energy_min": 0.0, "energy_max": EMAX}), nucleus=(24, 16, 40))
    fresh.contour._ctr_data = fs.contour._ctr_data     # the file's own nodes (numpy-version round-off of leggauss)
    fresh.concatFamData(full[:30], ["Strength"] + fs.xterms, [r[-1] for r in rows], [float(r[-2]) for r in rows])
    fresh._meta["Version"], fresh._meta["Interaction"] = "2.00", "SKOP"
    # the symmetric completion S(w*) = S(w)* reproduces the stored upper half exactly
    assert np.array_equal(fresh.cstr_df.values, full)
    fresh.writeStrengthOut(str(tmp_path))
    assert open(os.path.join(str(tmp_path), fs.file_txt)).read() == open(os.path.join(SOLN[op], fs.file_txt)).read()


def test_line_contours_and_settings():
    c = famContour("CONSTL", {"energy_min": 0.0, "energy_max": 3.0, "nr_points": 4, "half_width": 0.25})
    assert np.allclose(c.ctr_z, np.array([0, 1, 2, 3]) + 0.25j) and not c.closed and c.nr_compute == 4
    assert np.all(c.ctr_dzdt == 1) and np.all(c.theta == 0) and c.half_width == 0.25
    c = famContour("CONSTR", {"energy_min": 0.0, "energy_max": 1.0, "half_width": 0.5})
    assert np.allclose(c.ctr_z, np.array([0, 0.5, 1.0]) + 0.5j)
    with pytest.raises(KeyError):
        famContour("CIRCLE", {"no_such_key": 1})
    with pytest.raises(ValueError):
        famContour("SPIRAL")
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.0, "shift_imag": 0.1})
    assert c.nr_compute == 60
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 80.0, "nr_points": 9, "use_gauleg_ctr": False})
    assert c.nr_compute == 5 and abs(c.ctr_z[2].imag) <= 30 + 1e-12     # ellipse capped at max_height
    c.setHfbInterval({"E_gs": 1.2, "EQRPA_max": 7.0}, "-")
    assert (c.energy_min, c.energy_max) == (0.0, 7.0)
    # finite temperature (contour.py:184-198; checked against the reference module in the build container): the interval
    # starts at -30 MeV, and ends at +30 MeV for electron capture
    c.setHfbInterval({"E_gs": 1.2, "EQRPA_max": 7.0, "ft_active": True, "temperature": 0.8}, "-", shift=0.25)
    assert (c.energy_min, c.energy_max) == (-29.75, 7.25)
    c.setHfbInterval({"E_gs": 1.2, "EQRPA_max": 7.0, "ft_active": True, "temperature": 0.8}, "c")
    assert (c.energy_min, c.energy_max) == (-30.0, 30.0)


def test_odd_point_count_symmetry_fill():
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 4.0, "nr_points": 7})
    assert c.nr_compute == 4
    fs = famStrength("GT-", 0, c, nucleus=(24, 16, 40))
    s = (np.arange(4) + 1.0)[:, None] * np.array([[1 + 2j, 3 - 1j]])
    df = fs.concatFamStr(s, ["Strength", "GTxX"])
    assert list(df.columns) == ["Re(Strength)", "Im(Strength)", "Re(GTxX)", "Im(GTxX)"] and len(df) == 7
    assert np.array_equal(df["Re(Strength)"].values, [1, 2, 3, 4, 3, 2, 1])
    assert np.array_equal(df["Im(Strength)"].values, [2, 4, 6, 8, -6, -4, -2])


def test_patch_namelist():
    t = "&ext_field\n    beta_type = '-'\n    operator_name = 'GT'\n    operator_k = 0\n/\n"
    u = patch_namelist(t, operator_name="RS1", operator_k=1)
    assert "operator_name = 'RS1'" in u and "operator_k = 1" in u and "beta_type = '-'" in u
    with pytest.raises(KeyError):
        patch_namelist(t, nonexistent=1)


def beta_fixture():
    import json
    d = json.load(open(os.path.join(GOLDEN, "S40_SKOP_6sh", "beta_soln.json")))
    sf = {c: np.array(v["re"], float) + 1j * np.array(v["im"], float) for c, v in d["shape_factor"].items()}
    return sf, {c: float(v["rate"]) for c, v in d["rates"].items()}, {c: float(v["halflife"]) for c, v in d["rates"].items()}


@pytest.mark.parametrize("col,op,k", [("Allowed-GT_K=0", "GT-", 0), ("Allowed-GT_K=1", "GT-", 1)])
def test_integrated_rate_from_the_reference_shape_factor(col, op, k):
    """beta.out of the reference from its own shape factor columns: pins complex_quadrature / beta_rate; and the shape
    factor of an allowed channel is a smooth phase-space weight times the FAM strength stored in OP.out.ctr."""
    sf, rates, half = beta_fixture()
    fs = _fixture(op, k)
    rate, hl = beta_rate(fs.contour, sf[col])
    assert abs(rate - rates[col]) < 1e-13 * rates[col] and abs(hl - half[col]) < 1e-13 * half[col]
    w = sf[col] / fs.cstr_df["Strength"].values          # the reference's integration weights at the contour points
    assert np.all(np.isfinite(w)) and np.max(np.abs(w[:30] - np.conj(w[:29:-1]))) < 1e-9 * np.max(np.abs(w))
    with pytest.raises(ValueError):
        complex_quadrature("BOOLE", fs.contour, sf[col])
    line = famContour("CONSTL", {"energy_min": 0.0, "energy_max": 2.0, "nr_points": 5, "half_width": 0.1})
    assert abs(complex_quadrature("TRAP", line, np.ones(5) * 1j) - 2.0j) < 1e-15


def test_every_contour_type_against_the_reference_module():
    """All seven contour types, default and overridden settings: the arrays the reference's own famContour produces
    (tests/golden/contours.json, generated by importing pynfam/strength/contour.py -- tests/golden/make_contours.py)."""
    import json
    cases = json.load(open(os.path.join(GOLDEN, "contours.json")))["cases"]
    assert {c["name"] for c in cases} == {"CIRCLE", "CONSTL", "CONSTR", "FERMIS", "FERMIA", "EXP", "MONOMIAL"}
    for c in cases:
        k = famContour(c["name"], c["override"])
        tag = (c["name"], c["override"])
        assert (k.nr_points, k.nr_compute, k.closed, k.quadrature) == (c["nr_points"], c["nr_compute"], c["closed"], c["quadrature"]), tag
        assert k.name_and_int == c["name_and_int"], tag
        z = np.array(c["re"], float) + 1j * np.array(c["im"], float)
        dz = np.array(c["dzdt_re"], float) + 1j * np.array(c["dzdt_im"], float)
        tol = 1e-12 if c["name"] == "CIRCLE" else 0.0        # Gauss-Legendre nodes differ in the last bits between numpy builds
        assert np.max(np.abs(k.ctr_z - z)) <= tol * max(1.0, np.max(np.abs(z))), tag
        assert np.max(np.abs(k.ctr_dzdt - dz)) <= tol * max(1.0, np.max(np.abs(dz))), tag
