"""GPU parity at the sizes the numbers are quoted on (run with -m gpu on the B200 box), through the C ABI.

Known answers: the reference's OWN prebuilt pnfam_main.x run point by point, single-threaded, to convergence
(tests/golden/make_production.py; fixtures committed, nothing here reads /root/reference):
  * 162Gd 16 shells : 20 converged points of bench.py's contour sweep, incl. the four nearest-axis nodes (configs[4])
  * 162Gd 20 shells : one converged point for each of the 14 allowed + first-forbidden (operator, K) (configs[3])
  * 163Gd 16 shells : odd-A, blocked 5/2-[523] neutron, equal-filling P,Q quadrants (configs[2])
  * 162Gd 12 and 24 shells (configs[4]; 24 shells leaves the compile-time-specialised kernel instances)
  * the ill-conditioned 6-shell points of the reference's golden trees, re-run with the reference binary here
Bound: 1e-9 relative on S(omega) and the cross-terms (BASELINE.json north_star), identical iteration counts."""
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, stage_point, terminal_scatter  # noqa: F401
from pynfam_b200 import host

pytestmark = pytest.mark.gpu
TOL = 1e-9
# Points that need >= 25 Broyden steps are ill-conditioned: round-off differences between two correct implementations are
# amplified by ~1e7 and the stopping rule (max|dX| < 1e-7) may trigger one step earlier or later (the last steps still
# move S by ~1e-8).  The reference shows the same scatter against ITSELF: 8.1e-9 between its run in this build container
# and its own golden files (tests/golden/loose_points_6sh.json), and tests/golden/ref_thread_spread.json records its
# 1-thread vs 4-thread difference at the production-size points used here.  Bound for that class: 2.5 x 8.1e-9.
LOOSE_TOL = 2e-8
LOOSE_ITERS = 25


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pynfam_b200 import gpu as g
    return g


def omega_of(nml):
    return complex(float(re.search(r"real_eqrpa\s*=\s*(\S+)", nml).group(1)),
                   float(re.search(r"imag_eqrpa\s*=\s*(\S+)", nml).group(1)))


def stage(case, nml, wd, name):
    import shutil
    os.makedirs(wd, exist_ok=True)
    for f in os.listdir(os.path.join(GOLDEN, case)):
        if (f.startswith("hfbtho_") or f.endswith(".tbc")) and not os.path.isfile(os.path.join(wd, f)):
            shutil.copy(os.path.join(GOLDEN, case, f), wd)
    with open(os.path.join(wd, name), "w") as f:
        f.write(re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", nml))


def check_fixture(gpu, case, fname, wd, separable=True, slots=0, expect_efa=None, tol=TOL):
    """Every point of the fixture, one batched solve per operator; returns (points, worst relative error)."""
    path = os.path.join(GOLDEN, case, fname)
    if not os.path.isfile(path):
        pytest.skip("fixture %s/%s not generated" % (case, fname))
    pts = json.load(open(path))["points"]
    base, ctx, n, worst, worst_loose, n_loose, n_shift = None, None, 0, 0.0, 0.0, 0, 0
    for op, lst in pts.items():
        stage(case, lst[0]["namelist"], wd, op + ".in")
        p = host.Problem(wd, op + ".in", share_nucleus_with=base)
        if base is None:
            base, ctx = p, gpu.Context(p, separable=separable)
            assert ctx.separable == separable
            if expect_efa is not None:
                assert bool(p.iscalar("blo_active")) == expect_efa
        r = ctx.solve(p, omegas=[omega_of(pt["namelist"]) for pt in lst], slots=slots)
        for i, pt in enumerate(lst):
            assert pt["conv"], "fixture point did not converge in the reference"
            loose = pt["iters"] >= LOOSE_ITERS
            it_gpu, it_ref = int(r["iters"][i]), pt["iters"]
            assert int(r["conv"][i]) == 1
            rows = {k: complex(float(v[0]), float(v[1])) for k, v in pt["rows"].items()}
            s_ref = rows["Strength"]
            if it_gpu != it_ref:
                # The stopping rule max|dX| < eps triggered a step or two apart (only tolerated in the ill-conditioned class).
                # Stopped one step earlier: the state must equal the reference's OWN state at that iteration (its trace
                # prints 10 digits).  Otherwise: within the reference's own terminal movement of S (conftest.terminal_scatter).
                # (a point within 0.5 MeV of the real axis belongs to that class whatever its iteration count: at
                # Gd162 T=0.8 GT-K0 w = 6 + 0.25i the reference's residual reads 1.8e-7, 1.9e-7, 6.5e-8 over its last three steps)
                near_axis = abs(omega_of(pt["namelist"]).imag) < 0.5
                assert (loose or near_axis) and abs(it_gpu - it_ref) <= 2, (op, i, it_gpu, it_ref)
                tr = {t[0]: complex(t[3], t[4]) for t in pt["trace"]}
                err = abs(r["strength"][i, 0] - s_ref) / abs(s_ref)
                if it_gpu == it_ref - 1 and abs(r["strength"][i, 0] - tr[it_gpu]) / abs(s_ref) < LOOSE_TOL:
                    err = abs(r["strength"][i, 0] - tr[it_gpu]) / abs(s_ref)
                else:
                    assert err < LOOSE_TOL + 1.5 * terminal_scatter(pt), (case, op, i, it_gpu, it_ref, err, terminal_scatter(pt))
                worst_loose = max(worst_loose, err)
                n_shift += 1
            else:
                floor = 1e-6 * max(abs(v) for k, v in rows.items() if k != "Energy")
                for k, lab in enumerate(["Strength"] + r["labels"][1:]):
                    if lab in rows:
                        err = abs(r["strength"][i, k] - rows[lab]) / max(abs(rows[lab]), floor)
                        if loose:
                            worst_loose = max(worst_loose, err)
                        else:
                            worst = max(worst, err)
                        assert err < (max(tol, LOOSE_TOL) if loose else tol), (case, op, i, lab, err)
            n += 1
            n_loose += loose
    print("%s/%s: %d points: worst relative error %.2e on the %d well-conditioned ones (< 25 iterations), %.2e on the %d others "
          "(%d of them stopped one iteration apart from the reference)" % (case, fname, n, worst, n - n_loose, worst_loose, n_loose, n_shift))
    return n, worst


@pytest.mark.parametrize("separable", [True, False])
def test_gd162_16_shells_converged_sweep_points(gpu, tmp_path, separable):
    """The bench workload itself: 20 converged points of the contour sweep (18-54 iterations), nearest-axis nodes included."""
    n, _ = check_fixture(gpu, "Gd162_SKOP_16sh", "prod_points.json", str(tmp_path), separable=separable)
    assert n >= 16


def test_gd162_16_shells_through_few_slots(gpu, tmp_path):
    """Same points through 6 slots: admission on the device, results unchanged."""
    check_fixture(gpu, "Gd162_SKOP_16sh", "prod_points.json", str(tmp_path), slots=6)


@pytest.mark.parametrize("separable", [True, False])
def test_gd162_20_shells_all_operators_converged(gpu, tmp_path, separable):
    """configs[3]: F, GT, 0-, 1-, 2- operators (14 (operator, K)) at 20 shells, cross-terms by label."""
    n, _ = check_fixture(gpu, "Gd162_SKOP_20sh", "prod_points.json", str(tmp_path), separable=separable)
    assert n >= 8


@pytest.mark.parametrize("separable", [True, False])
def test_gd163_blocked_16_shells(gpu, tmp_path, separable):
    """configs[2] at its stated size: odd-A equal filling (8 amplitude vectors) on the 40-point-grid kernels."""
    n, _ = check_fixture(gpu, "Gd163_blocked_16sh", "points.json", str(tmp_path), separable=separable, expect_efa=True)
    assert n >= 4


@pytest.mark.parametrize("shells", [16, 20])
def test_gd162_finite_temperature_production_sizes(gpu, tmp_path, shells):
    """Finite temperature (T = 0.8 MeV) at the bench and north-star basis sizes: thermal occupations, 8 amplitude vectors
    (X, Y, P, Q) and T factors on the factorised kernels; HFB solutions and known answers from the reference's
    executables (tests/golden/make_production.py gd162_ft_16sh / gd162_ft_20sh)."""
    case = "Gd162_finiteT_%dsh" % shells
    path = os.path.join(GOLDEN, case, "points.json")
    if not os.path.isfile(path):
        pytest.skip("fixture %s not generated" % case)
    n, _ = check_fixture(gpu, case, "points.json", str(tmp_path), expect_efa=False)
    assert n >= 4
    stage(case, json.load(open(path))["points"]["GT-K0"][0]["namelist"], str(tmp_path / "p"), "x.in")
    assert host.Problem(str(tmp_path / "p"), "x.in").iscalar("ft_active") == 1


def test_gd162_16_shells_two_body_currents(gpu, tmp_path):
    """BASELINE configs[1] at the bench basis size: GT, RS0, P, PS0 with the closed-form (nuclear matter + LDA) two-body
    currents at 16 shells, known answers from the reference binary (tests/golden/make_production.py gd162_2bc_16sh)."""
    n, _ = check_fixture(gpu, "Gd162_SKOP_16sh", "tbc_points.json", str(tmp_path))
    assert n >= 5


def test_gd162_20_shells_sweep_points(gpu, tmp_path):
    """Six converged points of bench.py's contour sweep at 20 shells (18 to 54 iterations), the ones `bench.py --shells 20`
    checks its strengths against."""
    n, _ = check_fixture(gpu, "Gd162_SKOP_20sh", "sweep_points.json", str(tmp_path))
    assert n >= 6


def test_gd162_12_shells(gpu, tmp_path):
    check_fixture(gpu, "Gd162_SKOP_12sh", "points.json", str(tmp_path))


@pytest.mark.parametrize("separable", [True, False])
def test_gd162_24_shells(gpu, tmp_path, separable):
    """24 shells (N = 5850, nxy = 594 074): 13 n_z slots -> the generic (run-time stride) kernel instances, spin
    segments close to the 96-state limit of the projection accumulators."""
    check_fixture(gpu, "Gd162_SKOP_24sh", "points.json", str(tmp_path), separable=separable)


def test_ill_conditioned_points_against_the_reference_run_here(gpu, tmp_path):
    """Points that need >= 25 Broyden steps or lie within 0.5 MeV of the real axis amplify round-off: the fixture
    records that the reference binary, run in this project's build container, differs from the reference's OWN golden
    files by up to 8.1e-9 there (and takes 56 instead of 62, 41 instead of 44 iterations at two of them).  Measured
    here: this solver against the reference-run-here values.  Bounds: 1e-9 wherever reference-here reproduces
    reference-golden to 1e-10 and the iteration counts agree (the well-conditioned subset); everywhere else 2.5x the
    reference's own worst here-vs-golden spread."""
    d = json.load(open(os.path.join(GOLDEN, "loose_points_6sh.json")))["cases"]
    ref_spread = max(r["rel"] for lst in d.values() for r in lst)
    assert 1e-9 < ref_spread < 2e-8            # the recorded fact the relaxed bound rests on
    worst_all, worst_well, n_well, n_all = 0.0, 0.0, 0, 0
    for case, lst in d.items():
        allpts = json.load(open(os.path.join(GOLDEN, case, "points.json")))["points"]
        wd = str(tmp_path / case)
        base, ctx = None, None
        byop = {}
        for x in lst:
            byop.setdefault(x["name"], []).append(x)
        for op, rs in byop.items():
            nml = {p["point"]: p["namelist"] for p in allpts[op]}
            stage(case, nml[rs[0]["point"]], wd, op + ".in")
            p = host.Problem(wd, op + ".in", share_nucleus_with=base)
            if base is None:
                base, ctx = p, gpu.Context(p)
            r = ctx.solve(p, omegas=[omega_of(nml[x["point"]]) for x in rs])
            for i, x in enumerate(rs):
                here = complex(float(x["here"][0]), float(x["here"][1]))
                err = abs(r["strength"][i, 0] - here) / abs(here)
                assert int(r["conv"][i]) == 1
                worst_all = max(worst_all, err)
                n_all += 1
                well = x["rel"] < 1e-10 and x["here_iters"] == x["golden_iters"] == int(r["iters"][i])
                if well:
                    n_well += 1
                    worst_well = max(worst_well, err)
                    assert err < TOL, (case, op, x["point"], err)
                else:
                    assert err < 2.5 * ref_spread, (case, op, x["point"], err)
                    assert abs(int(r["iters"][i]) - x["here_iters"]) <= max(4, x["here_iters"] // 5)
    print("ill-conditioned 6-shell points: %d (well-conditioned subset %d): worst vs reference-here %.2e (subset %.2e); "
          "reference here-vs-golden %.2e" % (n_all, n_well, worst_all, worst_well, ref_spread))
    assert n_well >= 200


def test_closed_form_two_body_current_modes(gpu, tmp_path):
    """configs[1] beyond the .tbc-fed Gamow-Teller field: nuclear-matter / LDA two-body currents of GT (symmetric and
    asymmetric matter, 1BC+2BC and 2BC only), the (1 - correction) weight of RS0 / RS1 / RS2, the P and PS0 currents, with
    the corrected cross-term fields -- against the reference binary (tests/golden/make_2bc_modes.py)."""
    n, worst = check_fixture(gpu, "S40_2bc_modes", "points.json", str(tmp_path))
    assert n == 11
