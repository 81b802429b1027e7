"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libpnfam_b200.so (pynfam_b200.gpu is a thin ctypes view).  The checker is the CPU oracle (oracle/) and the
reference's golden vectors (tests/golden/); tolerance 1e-9 relative on the complex strength and cross-terms
as stated by BASELINE.json's north_star (observed agreement is ~1e-14)."""
import numpy as np
import pytest

from conftest import gold_rows, load_points, stage_point, terminal_scatter
from oracle import fam_oracle as fo
from pynfam_b200 import host

pytestmark = pytest.mark.gpu
TOL = 1e-9
# Ill-conditioned points (>= 25 Broyden steps or |Im omega| < 0.5): 2.5 x the spread the reference binary itself shows
# between its run in this build container and its own golden files at those points (8.1e-9, recorded in
# tests/golden/loose_points_6sh.json; measured against the reference-run-here in tests/test_gpu_production.py)
LOOSE_TOL = 2e-8


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pynfam_b200 import gpu as g
    return g


@pytest.fixture(params=["separable", "general"])
def mkctx(gpu, request):
    """Context factory for both kernel families: the sum-factorised ones (the model carries the separable factors
    of the HO basis -- the default) and the general-table ones."""
    sep = request.param == "separable"

    def make(problem):
        ctx = gpu.Context(problem, separable=sep)
        assert ctx.separable == sep
        return ctx
    return make


def _rel(a, b):
    return abs(a - b) / abs(b)


CASES = [
    ("S40_SKOP_6sh", "GT-K0", 10),
    ("S40_GT_All", "RS0-K0", 5),
    ("S40_GT_All", "P-K1", 3),
    ("S40_GT_All", "RS2-K2", 7),
    ("S40_GT_All", "R-K0", 1),
    ("S40_GT_All", "F-K0", 2),
    ("Gd162_GT_open_6sh", "GT-K1", 40),
    ("Gd162_1-_closed_6sh", "RS1-K1", 4),
    ("Gd162_0-_closed_6sh", "PS0-K0", 8),
    ("S40_All_GT2bc", "GT-K1", 12),      # two-body currents via .tbc
    ("Gd163_blocked_6sh", "GT-K0", 0),   # odd-A equal-filling: 8 amplitude vectors (X,Y,P,Q re/im)
    ("Gd163_blocked_6sh", "GT-K1", 0),
    ("Gd163_blocked_6sh", "RS1-K1", 0),
    ("Gd162_finiteT_6sh", "GT-K0", 0),   # finite temperature T = 0.8 MeV: thermal occupations, P,Q quadrants, T factors
    ("Gd162_finiteT_6sh", "GT-K1", 0),
    ("Gd162_finiteT_6sh", "F-K0", 0),
    ("Gd162_finiteT_6sh", "RS1-K1", 0),
    ("Gd162_finiteT_6sh", "RS0-K0", 0),
    ("Gd162_finiteT_6sh", "GT-K0", 1),   # interrupted at max_iter = 5
    ("S40_custom_interaction", "GT-K0", 0),   # couplings from a file (interaction_name = 'FILE:custom_edf.dat')
    ("S40_custom_interaction", "F-K0", 0),
]


@pytest.mark.parametrize("case,op,idx", CASES)
def test_trajectory_matches_oracle_and_golden(mkctx, case, op, idx, tmp_path):
    pt = stage_point(case, op, idx, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = mkctx(p)
    model = fo.model_from_problem(p)
    for mi in (1, 2, 4):
        so = fo.solver_from_problem(p, model)
        _, si, st = so.solve(mi, 1e-7)
        r = ctx.solve(p, max_iter=mi)
        assert abs(r["si"][0] - si) <= 1e-9 * si
        for k in range(len(st)):
            assert _rel(r["strength"][0, k], st[k]) < TOL, (mi, k)
    r = ctx.solve(p, want_trace=True)
    gold = gold_rows(pt)
    assert int(r["iters"][0]) == pt["iters"] and int(r["conv"][0]) == int(bool(pt["conv"]))
    for k, lab in enumerate(["Strength"] + r["labels"][1:]):
        if lab in gold:
            assert _rel(r["strength"][0, k], gold[lab]) < TOL, lab
    for (i, _, si_g, re_g, im_g) in pt["trace"]:
        t = r["trace"][0, i]
        assert abs(t[0] - si_g) < 6e-11 and abs(t[1] - re_g) < 6e-11 and abs(t[2] - im_g) < 6e-11


@pytest.mark.parametrize("case,op", [("S40_SKOP_6sh", "GT-K0"), ("S40_GT_All", "RS1-K1"), ("Gd162_GT_open_6sh", "GT-K0")])
def test_whole_contour_batched_against_golden(mkctx, case, op, tmp_path):
    """All omega points of one operator in ONE batched call (how the product is meant to be driven)."""
    pts = load_points(case)[op]
    stage_point(case, op, 0, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = mkctx(p)
    import re
    om = []
    for pt in pts:
        om.append(complex(float(re.search(r"real_eqrpa\s*=\s*(\S+)", pt["namelist"]).group(1)),
                          float(re.search(r"imag_eqrpa\s*=\s*(\S+)", pt["namelist"]).group(1))))
    r = ctx.solve(p, omegas=om)
    for i, pt in enumerate(pts):
        gold = gold_rows(pt)
        # Points that need >= 25 Broyden steps (close to the real axis / inside the dense part of the spectrum)
        # amplify round-off (the numpy oracle vs the golden file: 1e-15 for points 0-24 of S40 GT-K0, then 8e-10,
        # 4e-9 at points 26, 28): the reference
        # binary run HERE differs from its own golden file by 8.1e-9 at S40 GT-K0 point 28, and takes 26 instead
        # of 24 iterations at Gd162 GT-K0 point 43 (Im omega = 0.1) -- see DESIGN.md "Parity".  Such points are
        # held to the solver's own accuracy (eps = 1e-7); all others to 1e-9 and the exact iteration count.
        loose = abs(om[i].imag) < 0.5 or pt["iters"] >= 25
        if loose:
            assert int(r["conv"][i]) == 1 and abs(int(r["iters"][i]) - pt["iters"]) <= max(4, pt["iters"] // 5), i
        else:
            assert int(r["iters"][i]) == pt["iters"], i
        tol = LOOSE_TOL if loose else TOL
        for k, lab in enumerate(["Strength"] + r["labels"][1:]):
            if lab in gold:
                assert _rel(r["strength"][i, k], gold[lab]) < tol, (i, lab)


def test_calc_hamiltonian_entry_matches_oracle(mkctx, tmp_path):
    stage_point("Gd162_GT_open_6sh", "GT-K1", 40, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = mkctx(p)
    s = fo.solver_from_problem(p)
    s.iterate(0)
    s.iterate(1)
    order = [(11, 0), (11, 1), (12, 0), (12, 1), (22, 0), (22, 1), (21, 0), (21, 1)]
    ins = [(s.dRsp_im if c else s.dRsp_re).m[q].copy() for q, c in order]
    ref = [(s.dHsp_im if c else s.dHsp_re).m[q].copy() for q, c in order]
    outs = [r.copy() for r in ref]
    for o in outs:
        o.elem = np.zeros_like(o.elem)
    ctx.calc_hamiltonian(ins, outs)
    scale = max(np.abs(r.elem).max() for r in ref)
    for o, r in zip(outs, ref):
        assert np.abs(o.elem - r.elem).max() < 1e-12 * scale
    # linearity (size-independent property): H(2x - 3y) = 2 H(x) - 3 H(y)
    rng = np.random.default_rng(7)
    x = [b.copy() for b in ins]
    y = [b.copy() for b in ins]
    for b in y:
        b.elem = rng.standard_normal(len(b.elem))
    z = [b.copy() for b in ins]
    for bz, bx, by in zip(z, x, y):
        bz.elem = 2 * bx.elem - 3 * by.elem
    hx, hy, hz = ([r.copy() for r in ref] for _ in range(3))
    ctx.calc_hamiltonian(x, hx)
    ctx.calc_hamiltonian(y, hy)
    ctx.calc_hamiltonian(z, hz)
    sc = max(np.abs(b.elem).max() for b in hy)
    for a, b, c in zip(hx, hy, hz):
        assert np.abs(c.elem - (2 * a.elem - 3 * b.elem)).max() < 1e-12 * sc


def test_mixer_variants_match_oracle(gpu, tmp_path):
    """Broyden ring-buffer wrap-around (M=3), linear mixing (M=0) and no residual interaction (M<0)."""
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = gpu.Context(p)
    model = fo.model_from_problem(p)
    for M, mi in ((3, 12), (1, 6), (0, 8)):
        so = fo.FamSolver(model, p.i32("f_ir2c"), p.f64("f_elem"), [], True,
                          complex(p.scalar("real_eqrpa"), p.scalar("imag_eqrpa")), 1.0, M)
        _, si, st = so.solve(mi, 1e-7)
        r = ctx.solve(p, max_iter=mi, history=M)
        assert _rel(r["strength"][0, 0], st[0]) < TOL and abs(r["si"][0] - si) < 1e-9 * si, M
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path), name="none.in",
                patch=lambda s: s.replace("interaction_name = 'SKOP'", "interaction_name = 'NONE'"))
    pn = host.Problem(str(tmp_path), "none.in", share_nucleus_with=p)
    r = gpu.Context(pn).solve(pn)
    ref = complex(2.3502388218313224888E-01, -5.6592934893661954454E-02)   # live reference binary, SURVEY.md 8c
    assert int(r["iters"][0]) == 2 and r["si"][0] == 0.0 and _rel(r["strength"][0, 0], ref) < TOL


def test_symmetry_and_batch_independence(gpu, tmp_path):
    """S(omega*) = S(omega)* (the symmetry pynfam uses to fill half of a closed contour,
    pynfam/strength/fam_strength.py:292-338); a batched solve equals the same points solved one at a time."""
    stage_point("S40_GT_All", "RS1-K0", 3, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = gpu.Context(p)
    w = complex(p.scalar("real_eqrpa"), p.scalar("imag_eqrpa"))
    oms = [w, w.conjugate(), w + 1.5, w + 4.0 - 0.7j]
    rb = ctx.solve(p, omegas=oms)
    assert np.abs(rb["strength"][1] - rb["strength"][0].conj()).max() < 1e-12 * np.abs(rb["strength"][0]).max()
    for i, o in enumerate(oms):
        r1 = ctx.solve(p, omegas=[o])
        assert int(r1["iters"][0]) == int(rb["iters"][i])
        assert np.abs(r1["strength"][0] - rb["strength"][i]).max() <= 1e-13 * np.abs(rb["strength"][i]).max()


def test_gd162_16_shells_against_reference_binary(mkctx, tmp_path):
    """Full production size (N=1958, nghl=1600, nxy=103926): known answers produced by the reference's own
    pnfam_main.x (tests/golden/make_gd162_16sh.py)."""
    pts = load_points("Gd162_SKOP_16sh")
    ctx = None
    base = None
    for op, lst in pts.items():
        for i, pt in enumerate(lst):
            stage_point("Gd162_SKOP_16sh", op, i, str(tmp_path), name="%s_%d.in" % (op, i))
            p = host.Problem(str(tmp_path), "%s_%d.in" % (op, i), share_nucleus_with=base)
            if base is None:
                base = p
                ctx = mkctx(p)
            r = ctx.solve(p)
            gold = gold_rows(pt)
            assert int(r["iters"][0]) == pt["iters"]
            for k, lab in enumerate(["Strength"] + r["labels"][1:]):
                if lab in gold:
                    assert _rel(r["strength"][0, k], gold[lab]) < TOL, (op, i, lab)


def test_gd162_20_shells_multi_chunk_blocks(mkctx, tmp_path):
    """20 shells (N=3542): the largest blocks have spin segments longer than one 48-row chunk, which exercises the
    accumulation over a-chunks in the density kernel and several a-chunks per segment in the projection.  Known
    answers from the reference's pnfam_main.x at fixed iteration counts (tests/golden/make_gd162_20sh.py)."""
    pts = load_points("Gd162_SKOP_20sh")
    ctx = None
    base = None
    for op, lst in pts.items():
        for i, pt in enumerate(lst):
            stage_point("Gd162_SKOP_20sh", op, i, str(tmp_path), name="%s_%d.in" % (op, i))
            p = host.Problem(str(tmp_path), "%s_%d.in" % (op, i), share_nucleus_with=base)
            if base is None:
                base = p
                ctx = mkctx(p)
                assert (p.i32("num_spin_up") > 48).any() or ((p.i32("db") - p.i32("num_spin_up")) > 48).any()
            r = ctx.solve(p)
            gold = gold_rows(pt)
            assert int(r["iters"][0]) == pt["iters"]
            for k, lab in enumerate(["Strength"] + r["labels"][1:]):
                if lab in gold:
                    assert _rel(r["strength"][0, k], gold[lab]) < TOL, (op, i, lab)


def test_drop_in_executable_writes_the_reference_dat_contract(gpu, tmp_path):
    """pnfam_main.x <namelist> in a rundir, exactly as pynfam's fortProcess launches it
    (pynfam/fortran/fortran_utils.py:222-247); the .dat is parsed with the restated pnfamParser."""
    import os
    import subprocess
    from conftest import ROOT
    from oracle import refrun
    exe = os.path.join(ROOT, "pynfam_b200", "bin", "pnfam_main.x")
    pt = stage_point("S40_GT_All", "RS0-K0", 5, str(tmp_path), name="RS0-K0.in")
    r = subprocess.run([exe, "RS0-K0.in"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stderr.strip() == ""
    dat = refrun.parse_dat(open(str(tmp_path / "RS0-K0.dat")).read())
    assert dat["conv"] is True and dat["iters"] == pt["iters"] and dat["time_min"] is not None
    assert dat["version"].startswith("2.00")
    gold = gold_rows(pt)
    for lab, g in gold.items():
        assert _rel(dat["rows"][lab], g) < TOL if lab != "Energy" else abs(dat["rows"][lab] - g) < 1e-15
    assert refrun.parse_dat(r.stdout)["rows"].keys() == dat["rows"].keys()   # print_stdout = .true.
    for (i, lab, si_g, re_g, im_g), t in zip(pt["trace"], dat["trace"]):
        assert t[0] == i and abs(t[2] - si_g) < 6e-11 and abs(t[3] - re_g) < 6e-11


def test_contour_driver_writes_the_reference_strength_files(gpu, tmp_path):
    """one batched solve per operator -> OP.out + OP.out.ctr in pynfam's formats, checked against the
    reference's own files for the same nucleus and contour (tests/pynfam_test_S40 and tests/S40_GT_All, 000000/fam_soln)."""
    import os
    from conftest import GOLDEN
    from pynfam_b200.strength import famContour, famStrength, run_contours
    wd = str(tmp_path)
    stage_point("S40_SKOP_6sh", "GT-K0", 0, wd)
    contour = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036})
    res = run_contours(wd, "x.in", [("GT-", 0), ("GT-", 1), ("RS1-", 1)], contour)
    for fs in res:
        # GT from tests/pynfam_test_S40, RS1 from tests/S40_GT_All (same HFB solution; the older tree's RS1xP cross-term
        # predates the current operator definition)
        case = "S40_SKOP_6sh" if fs.bareop == "GT" else "S40_GT_All"
        soln = os.path.join(GOLDEN, case, "fam_soln")
        assert fs.meta["Conv"] == "Yes"
        ref = famStrength(fs.op, fs.k, "CIRCLE")
        ref.readCtrBinary(soln)
        got = famStrength(fs.op, fs.k, "CIRCLE")
        got.readCtrBinary(wd)                      # what betadecay / shapeFactor would read back
        assert got.nucleus == ref.nucleus == (24, 16, 40) and got.contour.nr_points == 60
        assert np.max(np.abs(got.contour.ctr_z - ref.contour.ctr_z)) < 1e-12
        a, b = got.cstr_df, ref.cstr_df
        z = ref.contour.ctr_z
        pts = load_points(case)[fs.opname]
        for lab in b.columns:                      # the 2023 fixture holds fewer cross-terms: compare by label
            assert lab in a.columns
            for i in range(60):
                j = i if i < 30 else 59 - i        # computed point this row mirrors
                loose = abs(z[j].imag) < 0.5 or pts[j]["iters"] >= 25
                assert _rel(a[lab].values[i], b[lab].values[i]) < (LOOSE_TOL if loose else TOL), (fs.opname, lab, i)
        if fs.bareop == "GT":
            # integrated beta-decay rate of the allowed channel (north-star: 1e-9 relative): the reference's own
            # phase-space weights at the contour points (its shape factor / its strength, tests/golden/make_beta.py)
            # applied to OUR strengths, integrated the way shapeFactor.calcBetaRates does, against its beta.out
            from pynfam_b200.strength import beta_rate
            from test_strength import beta_fixture
            sf, rates, _ = beta_fixture()
            col = "Allowed-GT_K=%d" % fs.k
            w = sf[col] / b["Strength"].values
            rate, _ = beta_rate(got.contour, w * a["Strength"].values)
            print("rate", col, rate, rates[col], abs(rate - rates[col]) / rates[col])
            assert abs(rate - rates[col]) < TOL * rates[col], (col, rate, rates[col])
        # the text summary parses the way pynfam re-reads it (strengthOutParser: header lines + whitespace table)
        lines = open(os.path.join(wd, fs.file_txt)).read().split("\n")
        assert lines[2] == "# All points converged: Yes" and lines[4] == "# Operator:             %s with K=%d" % (fs.op, fs.k)
        assert lines[8].split()[:5] == ["Theta", "Re(EQRPA)", "Im(EQRPA)", "Re(Strength)", "Im(Strength)"]
        row = lines[9 + 7].split()
        assert int(row[0]) == 7 and abs(float(row[4]) - b["Strength"].values[7].real) < 1e-9


def test_contour_executable_against_oracle(gpu, tmp_path):
    """contour_main.x, STR mode (exes/pnfam/contour_prog.f90 + contour_setup.f90): both 1+ operators on a 3-point line
    in one launch; the summary files hold what the oracle computes point by point, the per-point .dat files follow the
    pnfam_main.x contract."""
    import os
    import subprocess
    from conftest import ROOT
    from oracle import refrun
    from pynfam_b200.strength import patch_namelist
    wd = str(tmp_path)
    exe = os.path.join(ROOT, "pynfam_b200", "bin", "contour_main.x")
    stage_point("S40_SKOP_6sh", "GT-K0", 10, wd, name="pnfam_NAMELIST.dat")
    open(os.path.join(wd, "pnfam_CONTOUR.dat"), "w").write(
        "&ctr_general\n fam_mode = 'STR'\n fam_input_filename = 'pnfam_NAMELIST.dat'\n/\n"
        "&ctr_extfield\n operator_groups = '0+', '1+', '0-', '1-', '2-'\n operator_active = 0, 1, 0, 0, 0\n/\n"
        "&str_parameters\n energy_start = 1.0\n energy_step = 2.5\n nr_points = 3\n half_width = 0.75\n/\n")
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stderr.strip() == "", r.stdout[-500:]
    nml = open(os.path.join(wd, "pnfam_NAMELIST.dat")).read()
    for k in (0, 1):
        lines = open(os.path.join(wd, "GT-K%d.out" % k)).read().split("\n")
        assert lines[4] == "# Operator: GT with K = %d" % k and lines[5] == "# Gamma (half-width): .7500"
        assert lines[7].split()[:4] == ["#", "Conv", "Re(EQRPA)", "Re(Strength)"] and lines[7].startswith("# Conv")
        rows = [ln.split() for ln in lines[8:] if ln.strip()]
        assert len(rows) == 3
        open(os.path.join(wd, "o.in"), "w").write(patch_namelist(nml, operator_k=k))
        prob = host.Problem(wd, "o.in")
        model = fo.model_from_problem(prob)
        for i, row in enumerate(rows):
            w = complex(1.0 + 2.5 * i, 0.75)
            assert int(row[0]) == 1 and abs(float(row[1]) - w.real) < 1e-15
            dat = refrun.parse_dat(open(os.path.join(wd, "GT-K%d_%06d.dat" % (k, i))).read())
            assert dat["conv"] is True and abs(dat["rows"]["Energy"] - w) < 1e-15
            s = complex(float(row[2]), float(row[3]))
            assert _rel(s, dat["rows"]["Strength"]) < 1e-15
            so = fo.solver_from_problem(prob, model=model, omega=w)
            it, _, st = so.solve(300, 1e-7)
            tr = [t[-1] for t in so.trace]
            scatter = max(abs(tr[j] - tr[j - 1]) for j in range(len(tr) - 3, len(tr))) / abs(st[0])
            if dat["iters"] == it:
                # ill-conditioned class (>= 25 Broyden steps; 27 at w = 6 + 0.75i): within the oracle's own terminal movement of S
                assert _rel(s, st[0]) < (TOL if it < 25 else LOOSE_TOL + 1.5 * scatter), (k, i)
            else:
                # The stopping rule max|dX| < eps triggered a step or two apart.  Only legitimate when the oracle's residual at
                # the deciding iteration is borderline (within 15 % of eps: 1.04e-7 at step 21 of K = 1, w = 3.5 + 0.75i;
                # 9.2e-8 at step 27 of K = 0, w = 6 + 0.75i) or in the ill-conditioned class; the state is then compared at
                # EQUAL iteration numbers: the oracle is run for exactly the executable's count.
                res = {t[0]: t[2] for t in so.trace}
                first = min(it, dat["iters"])
                assert abs(it - dat["iters"]) <= 2 and (it >= 25 or abs(res[first] / 1e-7 - 1.0) < 0.15), (k, i, it, dat["iters"])
                _, _, st_eq = fo.solver_from_problem(prob, model=model, omega=w).solve(dat["iters"], 0.0)
                assert _rel(s, st_eq[0]) < LOOSE_TOL + 1.5 * scatter, (k, i, it, dat["iters"])


def _sharded_worker(rank, world, port, wd, dest):
    import os
    import torch.distributed as dist
    from pynfam_b200.strength import famContour, run_contours_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # both ranks share cuda:0 here; NCCL on a real box
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036, "nr_points": 20})
    run_contours_sharded(wd, "x.in", [("GT-", 0), ("GT-", 1), ("RS0-", 0)], c, dest=dest, dist=dist, device=0)
    dist.destroy_process_group()


def test_sharded_contour_driver_matches_the_single_process_run(gpu, tmp_path):
    """(operator, omega) tasks sharded over two ranks (whole points per rank, one all_reduce of the strengths) give the
    files of the single-process run: the strengths of a point do not depend on the batch it is solved in."""
    import os
    import socket
    import torch.multiprocessing as mp
    from pynfam_b200.strength import famContour, famStrength, run_contours
    wd, d1, d2 = str(tmp_path / "w"), str(tmp_path / "o1"), str(tmp_path / "o2")
    os.makedirs(d1), os.makedirs(d2)
    stage_point("S40_SKOP_6sh", "GT-K0", 0, wd)
    ops = [("GT-", 0), ("GT-", 1), ("RS0-", 0)]
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036, "nr_points": 20})
    run_contours(wd, "x.in", ops, c, dest=d1)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, wd, d2)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for op, k in ops:
        a, b = famStrength(op, k, "CIRCLE"), famStrength(op, k, "CIRCLE")
        a.readCtrBinary(d1)
        b.readCtrBinary(d2)
        x, y = a.cstr_df.values, b.cstr_df.values
        assert x.shape == y.shape and np.max(np.abs(x - y) / np.abs(x)) < 1e-12


def test_contour_driver_with_two_body_currents(gpu, tmp_path):
    """run_contours on the 2BC tree (tests/S40_All_GT2bc): every operator reads its own <OP>.tbc, so starting from the
    GT-K0 namelist the K=1 operator must pick GT-K1.tbc; all 30 computed points of both operators against the golden
    per-point results."""
    from pynfam_b200.strength import famContour, run_contours
    wd = str(tmp_path)
    stage_point("S40_All_GT2bc", "GT-K0", 0, wd)
    pts = {op: load_points("S40_All_GT2bc")[op] for op in ("GT-K0", "GT-K1")}
    contour = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036})
    res = run_contours(wd, "x.in", [("GT-", 0), ("GT-", 1)], contour)
    for fs in res:
        got = fs.cstr_df["Strength"].values
        for i, pt in enumerate(pts[fs.opname]):
            gold = gold_rows(pt)
            assert abs(contour.ctr_z[i] - gold["Energy"]) < 1e-12
            loose = abs(contour.ctr_z[i].imag) < 0.5 or pt["iters"] >= 25
            it_gpu, it_ref = int(fs.iters[i]), pt["iters"]
            if loose and 1 <= abs(it_gpu - it_ref) <= 2:
                # the stopping rule triggered a step or two apart on an ill-conditioned point: the result must lie within
                # the reference's own terminal movement of S (conftest.terminal_scatter)
                assert _rel(got[i], gold["Strength"]) < LOOSE_TOL + 1.5 * terminal_scatter(pt), (fs.opname, i, it_gpu, it_ref)
                continue
            assert it_gpu == it_ref or loose, (fs.opname, i, it_gpu, it_ref)
            assert _rel(got[i], gold["Strength"]) < (LOOSE_TOL if loose else TOL), (fs.opname, i)


def test_large_broyden_history_and_contour_mode_of_the_executable(gpu, tmp_path):
    """Any Broyden history size runs, as in the reference (its shipped run script uses broyden_history_size = 170): with
    more than 64 stored vectors the dots are staged in chunks and the Cholesky factor lives in global memory.  Compared
    with the oracle (numpy solve of the same system) after 80 iterations of a point that keeps iterating (eps = 0).
    Then the CONTOUR mode: contour_main.x fam_mode='CONTOUR' (a circle, absent from the reference) solves and lists
    every point."""
    import os
    import subprocess
    from conftest import ROOT
    wd = str(tmp_path)
    stage_point("S40_SKOP_6sh", "GT-K0", 10, wd, name="pnfam_NAMELIST.dat")
    p = host.Problem(wd, "pnfam_NAMELIST.dat")
    ctx = gpu.Context(p)
    model = fo.model_from_problem(p)
    for M, mi in ((170, 80), (70, 80), (64, 70)):
        so = fo.FamSolver(model, p.i32("f_ir2c"), p.f64("f_elem"), [], True,
                          complex(p.scalar("real_eqrpa"), p.scalar("imag_eqrpa")), 1.0, M)
        _, si, st = so.solve(mi, 0.0)
        r = ctx.solve(p, history=M, max_iter=mi, eps=0.0)
        assert int(r["iters"][0]) == mi and int(r["conv"][0]) == 0
        assert np.isfinite(r["si"][0]) and r["si"][0] < 1e-9 and _rel(r["strength"][0, 0], st[0]) < TOL, (M, r["si"][0], si)
    # a wrapped ring well beyond 64 on a point that is still far from convergence at every step is covered at fixed
    # iteration counts by test_mixer_variants_match_oracle; here the converged value must also equal the golden one
    r = ctx.solve(p, history=170)
    pt = load_points("S40_SKOP_6sh")["GT-K0"][10]
    assert int(r["iters"][0]) == pt["iters"] and _rel(r["strength"][0, 0], gold_rows(pt)["Strength"]) < TOL
    open(os.path.join(wd, "ctr.dat"), "w").write(
        "&ctr_general\n fam_mode = 'CONTOUR'\n fam_input_filename = 'pnfam_NAMELIST.dat'\n/\n"
        "&ctr_extfield\n operator_active = 0, 0, 0, 0, 0\n/\n"
        "&contour_parameters\n energy_min = 0.0\n energy_max = 8.0\n nr_points = 5\n shift_imag = 0.5\n/\n")
    exe = os.path.join(ROOT, "pynfam_b200", "bin", "contour_main.x")
    out = subprocess.run([exe, "ctr.dat"], cwd=wd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stderr.strip() == "", out.stdout[-400:]
    rows = [ln.split() for ln in open(os.path.join(wd, "GT-K0.out")).read().split("\n") if ln.strip() and not ln.startswith("#")]
    assert len(rows) == 5 and all(int(r_[0]) == 1 for r_ in rows)
    # equally spaced circle theta = pi .. 3 pi: Re(EQRPA) = 4 + 4 cos(theta)
    want = 4.0 + 4.0 * np.cos(np.pi + 2.0 * np.pi * np.arange(5) / 4.0)
    assert np.allclose([float(r_[1]) for r_ in rows], want, atol=1e-12)
    assert all(abs(complex(float(r_[2]), float(r_[3]))) > 0 for r_ in rows)


def test_full_beta_decay_chain_against_beta_out(gpu, tmp_path):
    """The whole flow the north-star names, on the GPU: all 14 (operator, K) of 40S on pynfam's contour (one batched
    solve each) -> phase space -> shape factor -> integrated rates, against the reference's beta.out
    (tests/S40_GT_All/000000/beta_soln).  The reference's own rate chain carries ~1e-9 of interpolation noise
    (tests/test_rates.py), the strengths of the near-axis points 2e-8."""
    import json
    import os
    from conftest import GOLDEN
    from pynfam_b200 import rates
    from pynfam_b200.strength import famContour, run_contours
    wd = str(tmp_path)
    stage_point("S40_GT_All", "GT-K0", 0, wd)
    ops = [("F-", 0), ("GT-", 0), ("GT-", 1), ("RS0-", 0), ("PS0-", 0), ("R-", 0), ("P-", 0), ("RS1-", 0), ("R-", 1), ("P-", 1),
           ("RS1-", 1), ("RS2-", 0), ("RS2-", 1), ("RS2-", 2)]
    contour = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036})
    fss = run_contours(wd, "x.in", ops, contour)
    gold = json.load(open(os.path.join(GOLDEN, "S40_GT_All", "beta_soln.json")))
    sf = rates.shapeFactor(fss, "-")
    sf.calcShapeFactor({k: float(v) for k, v in gold["hfb"].items()})
    df = sf.writeBetaOut(wd)
    total = float(gold["rates"]["Total"]["rate"])
    worst = 0.0
    for k, v in gold["rates"].items():
        err = abs(df.loc[k, "Rate(s^-1)"] - float(v["rate"])) / total
        worst = max(worst, err)
        assert err < 1e-8, (k, df.loc[k, "Rate(s^-1)"], v["rate"])
    print("rates vs beta.out: worst |d rate| / total = %.2e ; total %.16e vs %s" % (worst, df.loc["Total", "Rate(s^-1)"], gold["rates"]["Total"]["rate"]))
    assert abs(df.loc["Total", "Half-Life(s)"] / float(gold["rates"]["Total"]["halflife"]) - 1) < 1e-8
    assert os.path.isfile(os.path.join(wd, "beta.out"))


def test_continuous_batching_is_invisible_in_the_results(gpu, tmp_path):
    """Slots: 9 points through 2, 3 and 9 slots (points beyond the slot count are admitted on the device as running ones
    converge) give the same strengths, iteration counts and traces; an interrupted point (max_iter) frees its slot too."""
    pts = load_points("S40_GT_All")["RS1-K1"]
    stage_point("S40_GT_All", "RS1-K1", 0, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = gpu.Context(p)
    import re
    om = [complex(float(re.search(r"real_eqrpa\s*=\s*(\S+)", pt["namelist"]).group(1)),
                  float(re.search(r"imag_eqrpa\s*=\s*(\S+)", pt["namelist"]).group(1))) for pt in pts[4:28:3]] + [9.5 - 0.05j]
    ref = ctx.solve(p, omegas=om, slots=len(om), want_trace=True)
    assert ref["stats"]["batch_slots"] == len(om) and ref["stats"]["lock_steps"] >= int(ref["iters"].max())
    for S in (1, 2, 3):
        r = ctx.solve(p, omegas=om, slots=S, want_trace=True)
        assert r["stats"]["batch_slots"] == S and r["stats"]["iterations"] == ref["stats"]["iterations"]
        assert (r["iters"] == ref["iters"]).all() and (r["conv"] == ref["conv"]).all()
        assert np.abs(r["strength"] - ref["strength"]).max() <= 1e-13 * np.abs(ref["strength"]).max()
        assert np.abs(r["trace"][..., :3] - ref["trace"][..., :3]).max() <= 1e-12
    # interruption: max_iter below what most points need
    r = ctx.solve(p, omegas=om, slots=2, max_iter=12)
    r9 = ctx.solve(p, omegas=om, slots=9, max_iter=12)
    assert (r["iters"] == np.minimum(ref["iters"], 12)).all() and (r["conv"] == (ref["iters"] <= 12)).all()
    assert np.abs(r["strength"] - r9["strength"]).max() <= 1e-13 * np.abs(r9["strength"]).max()


def test_calc_hamiltonian_workspace_follows_the_block_structures(gpu, tmp_path):
    """The plug-in entry keeps its work space between calls (the reference calls it once per iteration with the same
    structures) and rebuilds it when the structures change: alternate two operators with different block maps."""
    res = {}
    ctx = None
    base = None
    for rnd in range(2):
        for case, op, idx in (("S40_GT_All", "GT-K1", 3), ("S40_GT_All", "RS2-K2", 7)):
            wd = str(tmp_path / ("%s_%d" % (op, rnd)))
            stage_point(case, op, idx, wd)
            p = host.Problem(wd, "x.in", share_nucleus_with=base)
            if base is None:
                base, ctx = p, gpu.Context(p)
            s = fo.solver_from_problem(p)
            s.iterate(0)
            s.iterate(1)
            order = [(11, 0), (11, 1), (12, 0), (12, 1), (22, 0), (22, 1), (21, 0), (21, 1)]
            ins = [(s.dRsp_im if c else s.dRsp_re).m[q].copy() for q, c in order]
            ref = [(s.dHsp_im if c else s.dHsp_re).m[q].copy() for q, c in order]
            outs = [r_.copy() for r_ in ref]
            for o in outs:
                o.elem = np.zeros_like(o.elem)
            ctx.calc_hamiltonian(ins, outs)
            scale = max(np.abs(r_.elem).max() for r_ in ref)
            assert max(np.abs(o.elem - r_.elem).max() for o, r_ in zip(outs, ref)) < 1e-12 * scale, (op, rnd)
            res.setdefault(op, []).append(np.concatenate([o.elem for o in outs]))
    for op, v in res.items():
        assert np.array_equal(v[0], v[1]), op     # same inputs, rebuilt work space: bit-identical


def test_empty_and_single_point_batches(gpu, tmp_path):
    """Edge cases of the batch: no omega point at all (nothing to do, no error) and a batch of one equal to the same
    point inside a larger batch."""
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    ctx = gpu.Context(p)
    r0 = ctx.solve(p, omegas=[])
    assert r0["strength"].shape == (0, 1) and len(r0["iters"]) == 0 and r0["stats"]["iterations"] == 0
    om = [0.58 - 2.4j, 3.0 + 1.0j, 7.0 - 0.7j]
    rb = ctx.solve(p, omegas=om)
    r1 = ctx.solve(p, omegas=om[1:2])
    assert int(r1["iters"][0]) == int(rb["iters"][1]) and _rel(r1["strength"][0, 0], rb["strength"][1, 0]) < 1e-13


def test_two_contexts_on_two_devices_in_one_process(gpu, tmp_path):
    """One process holding a context on each of two GPUs (per-device attribute / memory-pool caches, device guard): both
    solve the same points -- 16 shells, so that the kernels need more than 48 KB of dynamic shared memory on BOTH devices --
    alternately, the results are identical bit for bit and the caller's current device is left alone."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import os
    import re
    from conftest import GOLDEN
    pts = json.load(open(os.path.join(GOLDEN, "Gd162_SKOP_16sh", "prod_points.json")))["points"]["GT-K0"][:4]
    import shutil
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(GOLDEN, "Gd162_SKOP_16sh", f), str(tmp_path))
    (tmp_path / "x.in").write_text(pts[0]["namelist"])
    p = host.Problem(str(tmp_path), "x.in")
    om = [complex(float(re.search(r"real_eqrpa\s*=\s*(\S+)", q["namelist"]).group(1)),
                  float(re.search(r"imag_eqrpa\s*=\s*(\S+)", q["namelist"]).group(1))) for q in pts]
    torch.cuda.set_device(0)
    c0, c1 = gpu.Context(p, device=0), gpu.Context(p, device=1)
    r1 = c1.solve(p, omegas=om)
    assert torch.cuda.current_device() == 0
    r0 = c0.solve(p, omegas=om)
    r1b = c1.solve(p, omegas=om[:2])
    assert (r0["iters"] == r1["iters"]).all() and (r0["conv"] == 1).all()
    assert (r0["strength"] == r1["strength"]).all() and (r1b["strength"] == r1["strength"][:2]).all()
    for i, q in enumerate(pts):
        g = complex(float(q["rows"]["Strength"][0]), float(q["rows"]["Strength"][1]))
        assert _rel(r0["strength"][i, 0], g) < LOOSE_TOL
