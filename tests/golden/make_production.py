#!/usr/bin/env python
"""Production-size known answers from the reference's OWN prebuilt executables (oracle/_ref), farmed the way pynfam
farms them: one single-threaded pnfam_main.x process per omega point, `--jobs` side by side.

    python tests/golden/make_production.py <case> [--jobs 6]

Cases (BASELINE.json configs[2..4]; every point runs to convergence, max_iter = 300, eps = 1e-7, M = 50):
  gd162_16sh   162Gd 16 shells, GT- K=0: 20 stratified points of bench.py's contour sweep (incl. the four nearest-axis
               nodes of the first circle) -> Gd162_SKOP_16sh/prod_points.json
  gd162_20sh   162Gd 20 shells, one converged point for every one of the 14 allowed + first-forbidden (operator, K)
               on the 60-node CIRCLE contour of scripts/full_contour.py -> Gd162_SKOP_20sh/prod_points.json
  gd163_16sh   163Gd, 5/2-[523] blocked, 16 shells (configs[2]): hfbtho_main restarted from the even core with
               neutron_blocking = 5,-1,5,2,3, then GT K=0/1 and RS1 points -> Gd163_blocked_16sh/
  gd162_ft_16sh 162Gd at T = 0.8 MeV, 16 shells (finite-temperature HFB restarted from the zero-temperature solution), GT,
               F, RS0, RS1 points -> Gd162_finiteT_16sh/
  gd162_20sh_sweep six points of bench.py's contour sweep at 20 shells -> Gd162_SKOP_20sh/sweep_points.json
  gd162_2bc_16sh closed-form two-body currents (nuclear matter + LDA modes) at 16 shells -> Gd162_SKOP_16sh/tbc_points.json
  gd162_ft_20sh the same at 20 shells -> Gd162_finiteT_20sh/
  gd162_12sh / gd162_24sh   HFB ground state at 12 / 24 shells (same recipe as make_gd162_16sh.py) and GT K=0 sweep
               points -> Gd162_SKOP_{12,24}sh/
  loose_6sh    the ill-conditioned points of the reference's 6-shell golden trees (|Im omega| < 0.5 or >= 25
               iterations) re-run with the reference binary HERE, so "reference here vs reference golden" is a
               recorded number and not prose -> loose_points_6sh.json
Schema of every point: the one of make_golden.py (namelist, rows, iters, conv, trace, header) + sweep index.
"""
import argparse
import concurrent.futures as cf
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refrun  # noqa: E402
from make_gd162_16sh import FAM, SRC  # noqa: E402

ALL_OPS = [("F", 0), ("GT", 0), ("GT", 1), ("RS0", 0), ("PS0", 0), ("R", 0), ("R", 1), ("P", 0), ("P", 1),
           ("RS1", 0), ("RS1", 1), ("RS2", 0), ("RS2", 1), ("RS2", 2)]


def circle(npts, emin, emax):
    """pynfam CIRCLE contour (pynfam/strength/contour.py:212-283), all nodes -- same as bench.py:circle_contour."""
    x, _ = np.polynomial.legendre.leggauss(npts)
    theta = np.pi + (x + 1.0) * np.pi
    return 0.5 * (emin + emax) + 0.5 * (emax - emin) * np.exp(1j * theta)


def sweep(npts, nodes=64):
    out, g = [], 0
    while len(out) < npts:
        n = min(nodes, npts - len(out))
        out.extend(circle(n, 0.0, 10.0 + 0.25 * g))
        g += 1
    return np.array(out)


def hfb_ground_state(dst, shells, jobs):
    """Even-even 162Gd at `shells` HO shells, 40x40 grid, prolate start (recipe of make_gd162_16sh.py)."""
    os.makedirs(dst, exist_ok=True)
    if os.path.isfile(os.path.join(dst, "hfbtho_output.hel")):
        return
    wd = tempfile.mkdtemp()
    s = open(SRC).read()
    for a, b in (("number_of_shells = 6", "number_of_shells = %d" % shells), ("number_gauss = 20", "number_gauss = 40"),
                 ("number_laguerre = 20", "number_laguerre = 40"), ("number_legendre = 40", "number_legendre = 80"),
                 ("restart_file = 1", "restart_file = 2"), ("beta2_deformation = 0.0", "beta2_deformation = 0.3"),
                 ("basis_deformation = 0.0", "basis_deformation = 0.3")):
        assert a in s
        s = s.replace(a, b)
    open(os.path.join(wd, "hfbtho_NAMELIST.dat"), "w").write(s)
    out, t = refrun.run_hfbtho(wd, threads=jobs, timeout=4 * 3600)
    assert "iteration converged" in out, out[-2000:]
    print("hfbtho_main %d shells: %.0f s" % (shells, t), flush=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)


def hfb_blocked(dst, core, jobs):
    """163Gd with the 5/2-[523] neutron blocked, restarted from the even-even solution in `core`."""
    os.makedirs(dst, exist_ok=True)
    if os.path.isfile(os.path.join(dst, "hfbtho_output.hel")):
        return
    wd = tempfile.mkdtemp()
    s = open(os.path.join(core, "hfbtho_NAMELIST.dat")).read()
    s, n1 = re.subn(r"restart_file\s*=\s*-?\d+", "restart_file = -1", s)
    s, n2 = re.subn(r"neutron_blocking\s*=\s*0, 0, 0, 0, 0", "neutron_blocking = 5, -1, 5, 2, 3", s)
    assert n1 == 1 and n2 == 1
    open(os.path.join(wd, "hfbtho_NAMELIST.dat"), "w").write(s)
    shutil.copy(os.path.join(core, "hfbtho_output.hel"), wd)
    os.chmod(os.path.join(wd, "hfbtho_output.hel"), 0o644)
    out, t = refrun.run_hfbtho(wd, threads=jobs, timeout=4 * 3600)
    assert "iteration converged" in out, out[-3000:]
    print("hfbtho_main blocked: %.0f s" % t, flush=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)


def run_point(case_dir, op, k, w, max_iter, extra):
    wd = tempfile.mkdtemp()
    refrun.stage(wd, case_dir)
    name = "%s-K%d" % (op, k)
    nml = FAM.format(name=name, re=repr(float(w.real)), im=repr(float(w.imag)), op=op, k=k, max_iter=max_iter)
    extra = dict(extra)
    mode = extra.pop("two_body_current_mode", 0)
    if mode:
        nml = nml.replace("two_body_current_mode = 0", "two_body_current_mode = %d" % mode)
        extra["mode"] = mode
    open(os.path.join(wd, name + ".in"), "w").write(nml)
    dat, wall, out = refrun.run_pnfam(wd, name + ".in", threads=1, timeout=6 * 3600)
    shutil.rmtree(wd, ignore_errors=True)
    assert "Strength" in dat["rows"], out[-2000:]
    rec = {"namelist": nml, "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
           "iters": dat["iters"], "conv": dat["conv"],
           "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]], "header": dat["header"],
           "ref_wall_s": wall, "ref_threads": 1, "efa": "Odd-nucleus EFA active ..: Yes" in out}
    rec.update(extra)
    print(name, w, dat["rows"]["Strength"], dat["iters"], dat["conv"], "wall %.0fs" % wall, flush=True)
    return name, rec


def farm(case_dir, tasks, jobs, out_name, note):
    """tasks: [(op, k, omega, max_iter, extra dict)] -> case_dir/out_name"""
    points = {}
    path = os.path.join(case_dir, out_name)
    with cf.ThreadPoolExecutor(jobs) as ex:
        futs = [ex.submit(run_point, case_dir, *t) for t in tasks]
        for f in futs:
            name, rec = f.result()
            rec["point"] = "%06d" % len(points.get(name, []))
            points.setdefault(name, []).append(rec)
            json.dump({"source": note, "points": points}, open(path, "w"), indent=0)


NOTE = "reference's prebuilt hfbtho_main / pnfam_main.x (oracle/_ref), single-threaded, by tests/golden/make_production.py %s"


def gd162_16sh(jobs):
    d = os.path.join(HERE, "Gd162_SKOP_16sh")
    om = sweep(128)
    # first circle: the 4 nodes nearest to the real axis (both ends) + a stratified set of the lower half plane
    # (the upper half is its mirror image: S(conj w) = conj S(w) is a separate GPU property test); second circle: 4
    idx = [0, 63, 31, 32] + [2, 5, 8, 11, 14, 17, 20, 23, 26, 28, 29, 30] + [64 + 6, 64 + 18, 64 + 27, 64 + 40]
    tasks = [("GT", 0, om[i], 300, {"sweep_index": i}) for i in idx]
    farm(d, tasks, jobs, "prod_points.json", NOTE % "gd162_16sh")


def gd162_20sh(jobs):
    d = os.path.join(HERE, "Gd162_SKOP_20sh")
    c = circle(60, 0.0, 10.0)   # scripts/full_contour.py's contour; computed points are the first 30
    pick = [9, 14, 19, 23, 6, 12, 17, 21, 25, 8, 15, 20, 11, 27]
    tasks = [(op, k, c[pick[j]], 300, {"contour_index": pick[j]}) for j, (op, k) in enumerate(ALL_OPS)]
    farm(d, tasks, jobs, "prod_points.json", NOTE % "gd162_20sh")


def gd163_16sh(jobs):
    d = os.path.join(HERE, "Gd163_blocked_16sh")
    hfb_blocked(d, os.path.join(HERE, "Gd162_SKOP_16sh"), jobs)
    tasks = [("GT", 0, 1.0 + 0.5j, 300, {}), ("GT", 1, 3.0 + 1.0j, 300, {}), ("RS1", 1, 4.0 + 1.5j, 300, {}),
             ("GT", 0, 6.0 + 0.25j, 300, {}), ("GT", 0, 2.5 - 2.0j, 300, {}), ("RS0", 0, 5.0 + 3.0j, 300, {})]
    farm(d, tasks, jobs, "points.json", NOTE % "gd163_16sh")


def hfb_finite_temperature(dst, core, jobs, temperature):
    """162Gd at finite temperature, restarted from the zero-temperature solution in `core`."""
    os.makedirs(dst, exist_ok=True)
    if os.path.isfile(os.path.join(dst, "hfbtho_output.hel")):
        return
    wd = tempfile.mkdtemp()
    s = open(os.path.join(core, "hfbtho_NAMELIST.dat")).read()
    s, n1 = re.subn(r"restart_file\s*=\s*-?\d+", "restart_file = -1", s)
    s, n2 = re.subn(r"set_temperature\s*=\s*\.false\.", "set_temperature = .true.", s)
    s, n3 = re.subn(r"temperature\s*=\s*0\.0", "temperature = %r" % temperature, s)
    assert n1 == 1 and n2 == 1 and n3 == 1
    open(os.path.join(wd, "hfbtho_NAMELIST.dat"), "w").write(s)
    shutil.copy(os.path.join(core, "hfbtho_output.hel"), wd)
    os.chmod(os.path.join(wd, "hfbtho_output.hel"), 0o644)
    out, t = refrun.run_hfbtho(wd, threads=jobs, timeout=4 * 3600)
    assert "iteration converged" in out, out[-3000:]
    print("hfbtho_main T = %g: %.0f s" % (temperature, t), flush=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)


def gd162_finite_temperature_16sh(jobs):
    """162Gd at T = 0.8 MeV, 16 shells: thermal occupations, P,Q quadrants and T factors at the bench basis size."""
    d = os.path.join(HERE, "Gd162_finiteT_16sh")
    hfb_finite_temperature(d, os.path.join(HERE, "Gd162_SKOP_16sh"), jobs, 0.8)
    tasks = [("GT", 0, 1.0 + 0.5j, 300, {}), ("GT", 1, 3.0 + 1.0j, 300, {}), ("RS1", 1, 4.0 + 1.5j, 300, {}),
             ("GT", 0, 6.0 + 0.25j, 300, {}), ("F", 0, 2.5 - 2.0j, 300, {}), ("RS0", 0, 5.0 + 3.0j, 300, {})]
    farm(d, tasks, jobs, "points.json", NOTE % "gd162_ft_16sh")


def gd162_20sh_sweep(jobs):
    """Points of bench.py's contour sweep at 20 shells (bench.py --shells 20 compares its strengths with them outside the
    timed region) -> Gd162_SKOP_20sh/sweep_points.json"""
    d = os.path.join(HERE, "Gd162_SKOP_20sh")
    om = sweep(64)
    tasks = [("GT", 0, om[i], 300, {"sweep_index": i}) for i in (2, 11, 20, 26, 29, 31)]
    farm(d, tasks, jobs, "sweep_points.json", NOTE % "gd162_20sh_sweep")


def gd162_2bc_16sh(jobs):
    """BASELINE configs[1] at the bench basis size: operators with the closed-form (nuclear matter + LDA) two-body
    currents at 16 shells -- the mode digits of tests/golden/make_2bc_modes.py."""
    d = os.path.join(HERE, "Gd162_SKOP_16sh")
    tasks = [("GT", 0, 2.0 + 1.0j, 300, {"two_body_current_mode": 131100}), ("GT", 1, 4.0 + 1.5j, 300, {"two_body_current_mode": 121100}),
             ("RS0", 0, 5.0 + 2.0j, 300, {"two_body_current_mode": 121211}), ("P", 0, 3.0 + 1.0j, 300, {"two_body_current_mode": 221110}),
             ("PS0", 0, 6.0 + 2.5j, 300, {"two_body_current_mode": 121101})]
    farm(d, tasks, jobs, "tbc_points.json", NOTE % "gd162_2bc_16sh")


def gd162_finite_temperature_20sh(jobs):
    """162Gd at T = 0.8 MeV, 20 shells (the basis of the north-star run)."""
    d = os.path.join(HERE, "Gd162_finiteT_20sh")
    hfb_finite_temperature(d, os.path.join(HERE, "Gd162_SKOP_20sh"), jobs, 0.8)
    tasks = [("GT", 0, 1.0 + 0.5j, 300, {}), ("GT", 1, 3.0 + 1.0j, 300, {}), ("RS1", 1, 4.0 + 1.5j, 300, {}), ("RS0", 0, 5.0 + 3.0j, 300, {})]
    farm(d, tasks, jobs, "points.json", NOTE % "gd162_ft_20sh")


def gd162_small_large(shells, jobs, idx):
    d = os.path.join(HERE, "Gd162_SKOP_%dsh" % shells)
    hfb_ground_state(d, shells, jobs)
    om = sweep(64)
    tasks = [("GT", 0, om[i], 300, {"sweep_index": i}) for i in idx]
    farm(d, tasks, jobs, "points.json", NOTE % ("gd162_%dsh" % shells))


def loose_6sh(jobs):
    """Re-run, with the reference binary here, every golden 6-shell point the GPU tests hold to the relaxed bound."""
    out = {}
    for case in ("S40_SKOP_6sh", "S40_GT_All", "Gd162_GT_open_6sh", "Gd162_1-_closed_6sh", "Gd162_0-_closed_6sh"):
        cd = os.path.join(HERE, case)
        pts = json.load(open(os.path.join(cd, "points.json")))["points"]
        todo = []
        for name, lst in pts.items():
            for p in lst:
                im = float(p["rows"]["Energy"][1])
                if p["conv"] and (abs(im) < 0.5 or p["iters"] >= 25):
                    todo.append((name, p))

        def one(item):
            name, p = item
            wd = tempfile.mkdtemp()
            refrun.stage(wd, cd)
            nml = re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", p["namelist"])
            open(os.path.join(wd, name + ".in"), "w").write(nml)
            dat, wall, o = refrun.run_pnfam(wd, name + ".in", threads=1)
            shutil.rmtree(wd, ignore_errors=True)
            g = complex(float(p["rows"]["Strength"][0]), float(p["rows"]["Strength"][1]))
            h = dat["rows"]["Strength"]
            return {"case": case, "name": name, "point": p["point"], "golden_iters": p["iters"], "here_iters": dat["iters"],
                    "golden": [repr(g.real), repr(g.imag)], "here": [repr(h.real), repr(h.imag)],
                    "rows_here": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
                    "rel": abs(h - g) / abs(g)}
        with cf.ThreadPoolExecutor(jobs) as ex:
            res = list(ex.map(one, todo))
        out[case] = res
        print(case, len(res), "points, max rel here-vs-golden %.2e" % max([r["rel"] for r in res] or [0]), flush=True)
        json.dump({"source": NOTE % "loose_6sh", "cases": out}, open(os.path.join(HERE, "loose_points_6sh.json"), "w"), indent=0)


def thread_spread(threads):
    """Reference vs itself at production size: every fixture point that needs >= 25 iterations is run again with
    `threads` OpenMP/OpenBLAS threads (a different summation order inside its dgemm calls) and the difference to the
    single-threaded fixture value is recorded -> tests/golden/ref_thread_spread.json"""
    out = []
    for case, fname in (("Gd162_SKOP_12sh", "points.json"), ("Gd162_SKOP_16sh", "prod_points.json"), ("Gd163_blocked_16sh", "points.json"),
                        ("Gd162_SKOP_20sh", "prod_points.json"), ("Gd162_SKOP_24sh", "points.json")):
        cd = os.path.join(HERE, case)
        pts = json.load(open(os.path.join(cd, fname)))["points"]
        for name, lst in pts.items():
            for i, p in enumerate(lst):
                if p["iters"] < 25:
                    continue
                wd = tempfile.mkdtemp()
                refrun.stage(wd, cd)
                open(os.path.join(wd, name + ".in"), "w").write(p["namelist"])
                dat, wall, o = refrun.run_pnfam(wd, name + ".in", threads=threads, timeout=6 * 3600)
                shutil.rmtree(wd, ignore_errors=True)
                g = complex(float(p["rows"]["Strength"][0]), float(p["rows"]["Strength"][1]))
                h = dat["rows"]["Strength"]
                rec = {"case": case, "file": fname, "name": name, "index": i, "iters_1thread": p["iters"], "iters_threads": dat["iters"],
                       "threads": threads, "rel": abs(h - g) / abs(g), "strength_threads": [repr(h.real), repr(h.imag)]}
                out.append(rec)
                print(rec, flush=True)
                json.dump({"source": NOTE % ("thread_spread %d" % threads), "points": out},
                          open(os.path.join(HERE, "ref_thread_spread.json"), "w"), indent=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--jobs", type=int, default=6)
    a = ap.parse_args()
    if a.case == "gd162_16sh":
        gd162_16sh(a.jobs)
    elif a.case == "gd162_20sh":
        gd162_20sh(a.jobs)
    elif a.case == "gd163_16sh":
        gd163_16sh(a.jobs)
    elif a.case == "gd162_ft_16sh":
        gd162_finite_temperature_16sh(a.jobs)
    elif a.case == "gd162_20sh_sweep":
        gd162_20sh_sweep(a.jobs)
    elif a.case == "gd162_2bc_16sh":
        gd162_2bc_16sh(a.jobs)
    elif a.case == "gd162_ft_20sh":
        gd162_finite_temperature_20sh(a.jobs)
    elif a.case == "gd162_12sh":
        gd162_small_large(12, a.jobs, [0, 63, 4, 10, 16, 22, 27, 30, 31, 32])
    elif a.case == "gd162_24sh":
        gd162_small_large(24, a.jobs, [3, 9, 15, 21, 26, 29])
    elif a.case == "loose_6sh":
        loose_6sh(a.jobs)
    elif a.case == "thread_spread":
        thread_spread(a.jobs)
    else:
        raise SystemExit("unknown case")


if __name__ == "__main__":
    main()
