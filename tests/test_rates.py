"""Beta-decay rate chain (pynfam_b200/rates.py) against the reference's own outputs for 40S (tests/S40_GT_All, copied by
tests/golden/make_rates.py): phase-space integrals f1..f6 on the complex contour, every shape-factor column, every row
of beta.out -- all from the reference's 14 OP.out.ctr strength files.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from pynfam_b200 import rates
from pynfam_b200.strength import famStrength

CASE = os.path.join(GOLDEN, "S40_GT_All")
OPS = [("F-", 0), ("GT-", 0), ("GT-", 1), ("RS0-", 0), ("PS0-", 0), ("R-", 0), ("R-", 1), ("P-", 0), ("P-", 1),
       ("RS1-", 0), ("RS1-", 1), ("RS2-", 0), ("RS2-", 1), ("RS2-", 2)]


def gold():
    d = json.load(open(os.path.join(CASE, "beta_soln.json")))
    cx = lambda t: {k: np.array(v["re"], float) + 1j * np.array(v["im"], float) for k, v in t.items()}
    return d, cx(d["phase_space"]), cx(d["shape_factor"])


def reference_strengths():
    out = []
    for op, k in OPS:
        fs = famStrength(op, k, "CIRCLE")
        fs.readCtrBinary(os.path.join(CASE, "fam_soln"))
        fs.contour._settings.update(energy_min=0.0, energy_max=10.476036)
        out.append(fs)
    return out


@pytest.fixture(scope="module")
def sf():
    d, _, _ = gold()
    s = rates.shapeFactor(reference_strengths(), "-")
    s.updateSettings({"GA": d["settings"]["GA"], "GV": d["settings"]["GV"]})
    s.calcShapeFactor({k: float(v) for k, v in d["hfb"].items()})
    return s


def test_phase_space_on_the_contour(sf):
    """f_n(W0) continued to the complex contour by the Thiele interpolant: the reference's phasespace_{re,im}.out."""
    _, ps, _ = gold()
    for n in range(1, 7):
        a, b = sf.ps_df["f%d" % n], ps["f%d" % n]
        # 20-point Thiele interpolants are ill-conditioned: different numpy / scipy builds (nodes, loggamma) move the
        # continued values by ~1e-10 of the scale (observed 2e-11 .. 5e-10); pointwise relative agreement is 1e-14
        # except next to W0 = 1 where f_n -> 0
        assert np.max(np.abs(a - b)) < 2e-9 * np.max(np.abs(b)), n
        assert np.median(np.abs(a - b) / np.abs(b)) < 1e-9, n
    m = sf.sf_metadict
    assert (m["Zi"], m["Zf"], m["A"]) == (16, 17, 40) and abs(m["W0_max"] - 16.07022) < 5e-6 and abs(m["Radius"] - 0.01063) < 5e-6


def test_shape_factor_columns(sf):
    _, _, want = gold()
    assert set(want) == set(sf.betaout_keys)
    for k in sf.betaout_keys:
        scale = np.max(np.abs(want[k]))
        assert np.max(np.abs(sf.sf_df[k] - want[k])) < 2e-9 * scale, k


def test_rates_and_half_lives_of_beta_out(sf):
    d, _, _ = gold()
    df = sf.calcBetaRates()
    assert list(df.index) == sf.betaout_keys
    for k, v in d["rates"].items():
        # the interpolation noise above, integrated: 1e-9 of the total rate (observed: Total 1.1e-9, J=2 5e-12)
        assert abs(df.loc[k, "Rate(s^-1)"] - float(v["rate"])) < 3e-9 * float(d["rates"]["Total"]["rate"]), k
    for k in ("Total", "Total-Allowed", "Total-GT", "Total-Forbidden", "Allowed-GT_K=1", "Forbidden-J=2"):
        assert abs(df.loc[k, "Half-Life(s)"] / float(d["rates"][k]["halflife"]) - 1) < 3e-9, k
    assert abs(df.loc["Total", "Half-Life(s)"] - 4.7255073660069371) < 5e-8


def test_fermi_function_and_interpolant_basics():
    # non-relativistic limit: F_0 L_0 -> 2 pi y / (1 - exp(-2 pi y)) for small alpha Z (point nucleus)
    w = np.array([1.5, 3.0, 10.0])
    p = np.sqrt(w * w - 1)
    y = rates.ALPHA * 2 * w / p
    assert np.allclose(rates.Fermi(0, 2, 4, w) * rates.L0(2), 2 * np.pi * y / (1 - np.exp(-2 * np.pi * y)), rtol=2e-3)
    assert rates.Fermi(0, 20, 50, np.array([1.0, 0.5]))[0] == 0.0 and rates.lambda_ke(2, 20, 50, np.array([1.0]))[0] == 0.0
    x = np.linspace(0.0, 2.0, 9)
    t = rates.thieleInterpolator(x, 1.0 / (1.0 + x * x))
    assert abs(t(0.7 + 0.3j) - 1.0 / (1.0 + (0.7 + 0.3j) ** 2)) < 1e-9      # a rational function is reproduced off the axis
    ps = rates.phaseSpace("-")
    with pytest.raises(ValueError):
        ps.calcPsi(2, 17, 40, np.array([2.0 + 1.0j]))
    with pytest.raises(KeyError):
        ps.updateSettings({"nope": 1})
    with pytest.raises(NotImplementedError):
        rates.phaseSpace("c")
    f2 = ps.calcPsi(2, 17, 40, np.array([1.0, 5.0, 10.0]))
    assert f2[0] == 0.0 and f2[1] > 0 and f2[2] > f2[1]


def test_beta_out_file_layout(sf, tmp_path):
    """beta.out as pynfam writes it: the summary header of the reference's file (all lines but version / date) and the
    same rows; parsed back, the numbers are the rates."""
    df = sf.writeBetaOut(str(tmp_path))
    got = open(os.path.join(str(tmp_path), "beta.out")).read().split("\n")
    ref_header = [
        "# Nuclear Beta Decay Rates and Half-Lives", None, None, "# Summary Data:",
        "#   FAM_ctr     = CIRCLE on (0.00e+00, 1.05e+01)      , temper      = 0.00000   ",
        "#   beta_type   = -         , quadratr    = GAUSS     , Half_Width  = N/A       , screening   = N/A       ",
        "#   psi_approx  = RATINT    , psi_glpts   = 15        , ratint_pts  = 20        ",
        "#   Zi          = 16        , A           = 40        , Zf          = 17        , HFB_Qval    = 7.70087   ",
        "#   FAM_Qval    = N/A       , EQRPAmax    = 10.47604  , E_1stPeak   = N/A       , |gA|/gV     = 1.27000   ",
        "#   gA          = -1.27000  , gV          = 1.00000   , M_nucleon   = 939.00000 , alpha*Z     = 0.12405   ",
        "#   Radius      = 0.01063   , alpha*Z/2R  = 5.83646   , W0_max      = 16.07022  , W0*R        = 0.18096   ",
        "#"]
    for a, b in zip(got, ref_header):
        if b is not None:
            assert a == b
    assert got[12].split() == ["Rate(s^-1)", "Half-Life(s)"]
    rows = [ln.split() for ln in got[13:] if ln.strip()]
    assert [r[0] for r in rows] == sf.betaout_keys
    assert abs(float(rows[0][1]) - df.loc["Total", "Rate(s^-1)"]) < 1e-18 and len(rows[0][1]) == 22


def test_heavy_deformed_nucleus_gamow_teller_only():
    """162Gd (Z = 64, strong Coulomb distortion, small Q value), Gamow-Teller strengths only: phase space, shape
    factors and the rates of the reference's beta.out; the absent forbidden operators contribute exactly zero."""
    case = os.path.join(GOLDEN, "Gd162_GT_closed_6sh")
    d = json.load(open(os.path.join(case, "beta_soln.json")))
    cx = lambda t: {k: np.array(v["re"], float) + 1j * np.array(v["im"], float) for k, v in t.items()}
    ps, want = cx(d["phase_space"]), cx(d["shape_factor"])
    fss = []
    for k in (0, 1):
        fs = famStrength("GT-", k, "CIRCLE")
        fs.readCtrBinary(os.path.join(case, "fam_soln"))
        fs.contour._settings.update(energy_min=0.0, energy_max=d["settings"]["energy_max"])
        fss.append(fs)
    assert fss[0].nucleus == (98, 64, 162)
    s = rates.shapeFactor(fss, "-")
    s.calcShapeFactor({k: float(v) for k, v in d["hfb"].items()})
    for n in range(1, 7):
        a, b = s.ps_df["f%d" % n], ps["f%d" % n]
        assert np.max(np.abs(a - b)) < 2e-9 * np.max(np.abs(b)), n
    for k in ("Total", "Allowed-GT_K=0", "Allowed-GT_K=1"):
        assert np.max(np.abs(s.sf_df[k] - want[k])) < 2e-9 * np.max(np.abs(want[k])), k
    df = s.calcBetaRates()
    total = float(d["rates"]["Total"]["rate"])
    for k, v in d["rates"].items():
        assert abs(df.loc[k, "Rate(s^-1)"] - float(v["rate"])) < 3e-9 * total, k
    assert df.loc["Total-Forbidden", "Rate(s^-1)"] == 0.0 and np.isinf(df.loc["Total-Forbidden", "Half-Life(s)"])
    assert abs(df.loc["Total", "Half-Life(s)"] - 13.567947034686856) < 1e-7


def test_phase_space_functions_against_the_reference_module():
    """Fermi functions F_0 / F_1, lambda_2, real-axis integrals and the Thiele continuation for beta-minus and beta-plus
    (negative Z, Rose screening): values produced by importing the reference's phase_space.py
    (tests/golden/make_phase_space.py)."""
    g = json.load(open(os.path.join(GOLDEN, "phase_space.json")))
    w = np.array(g["w"])
    flt = lambda v: np.array(v, float)
    for r in g["fermi"]:
        got = rates.Fermi(r["F"], r["Z"], r["A"], w.copy(), r["sc"])
        assert np.allclose(got, flt(r["val"]), rtol=1e-13, atol=0), (r["F"], r["Z"], r["sc"])
    for r in g["lambda2"]:
        assert np.allclose(rates.lambda_ke(2, r["Z"], r["A"], w.copy(), r["sc"]), flt(r["val"]), rtol=1e-13, atol=0), r["Z"]
    for r in g["calcPsi"]:
        p = rates.phaseSpace(r["beta"])
        got = p.calcPsi(r["n"], r["Z"], r["A"], np.array(r["w0"]), sc=r["sc"])
        assert np.allclose(got, flt(r["val"]), rtol=1e-12, atol=0), (r["beta"], r["Z"], r["n"])
    for r in g["psiFct"]:
        p = rates.phaseSpace(r["beta"])
        z = np.array(r["re_z"]) + 1j * np.array(r["im_z"])
        got = p.psiFct(r["n"], r["Z"], r["A"], r["eqrpamax"], r["eqrpamin"])(z)
        want = flt(r["re"]) + 1j * flt(r["im"])
        # the 20-point continued fraction amplifies last-bit differences of its inputs (the real-axis integrals above
        # agree to 1e-12) where f_n is small: compare on the scale of the function (worst case 3e-9, beta+ Z = 63)
        assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), (r["beta"], r["Z"], r["n"])


def test_fermi_function_against_the_landolt_boernstein_table_of_the_reference_unit_test():
    """exes/pnfam/tests/modules/fermi_test.f90 (f0_us_test, lambda2_us_test): the reference's own unit test compares the
    unscreened Fermi function F0 and lambda_2 with the Landolt-Boernstein tables for Z = 32 (A = 72) and Z = 84 (A = 210) at
    eleven momenta p / m_e, with cut-offs of 0.1 % and 35 %.  Same table, same cut-offs (measured: 0.018 % and 25 %)."""
    import numpy as np
    from pynfam_b200 import rates
    p = np.array([0.1, 0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 10.0, 20.0, 25.0, 30.0])
    table = {   # (Z, A): (F0, lambda_2), fermi_test.f90:365-388
        (32, 72): ([2.1611e1, 4.8821, 3.2982, 2.7288, 2.5778, 2.5051, 2.4594, 2.3476, 2.2546, 2.2266, 2.2039],
                   [5.5523, 1.0796, 0.9524, 0.9344, 0.9396, 0.9464, 0.9531, 0.9801, 1.0199, 1.0367, 1.0525]),
        (84, 210): ([3.8164e2, 8.0476e1, 4.5386e1, 2.9197e1, 2.3706e1, 2.0707e1, 1.8726e1, 1.3865e1, 1.0340e1, 9.4135, 8.7180],
                    [15.6186, 1.1168, 0.6928, 0.6440, 0.6798, 0.7222, 0.7637, 0.9429, 1.2198, 1.3340, 1.4352]),
    }
    w = np.sqrt(p * p + 1.0)
    for (Z, A), (f0, l2) in table.items():
        assert np.max(np.abs(rates.Fermi(0, Z, A, w.copy(), False) / np.array(f0) - 1.0)) < 1e-3
        assert np.max(np.abs(rates.lambda_ke(2, Z, A, w.copy(), False) / np.array(l2) - 1.0)) < 0.35


def test_phase_space_integrals_against_the_mathematica_values_of_the_reference_unit_test():
    """exes/pnfam/tests/modules/phasespace_test.f90 (real_psi_test): f_1 .. f_6 for Z_final = 57, A = 174,
    W_max = 27.766732 against Mathematica, cut-off 5e-4 there (measured here: <= 4.8e-6)."""
    import numpy as np
    from pynfam_b200 import rates
    mathematica = [2.209608199614178e5, 2.577925271039772e6, 3.470883516733893e7, 5.398097056657456e8,
                   5.998661241258061e8, 5.178321508233396e8]
    ps = rates.phaseSpace("-")
    for n, ref in enumerate(mathematica, start=1):
        val = float(np.ravel(ps.calcPsi(n, 57, 174, np.array([27.766732])))[0])
        assert abs(val / ref - 1.0) < 5e-4
        assert abs(val / ref - 1.0) < 1e-5
