#!/usr/bin/env python
"""Generate tests/golden/Gd162_finiteT_6sh/: a finite-temperature HFB solution (T = 0.8 MeV) and the FAM known answers on
top of it (thermal quasiparticle occupations: P,Q quadrants and T factors active, pnfam_setup.f90:323-331,
pnfam_solver.f90:248-254, 440-457).

The reference ships no finite-temperature fixture.  Made with its own prebuilt executables (oracle/_ref):
  1. hfbtho_main restarted from the 162Gd solution of tests/golden/Gd162_GT_open_6sh with set_temperature = .true.,
     temperature = 0.8, converged;
  2. pnfam_main.x known answers for a few (operator, omega) points.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refrun  # noqa: E402
from make_gd162_16sh import FAM  # noqa: E402

POINTS = [("GT", 0, 1.0 + 0.5j, 300), ("GT", 1, 3.0 + 1.0j, 300), ("F", 0, 2.0 + 0.75j, 300), ("RS1", 1, 4.0 + 1.5j, 300),
          ("RS0", 0, 5.0 + 2.0j, 300), ("GT", 0, 6.0 + 0.25j, 5)]


def main():
    src = os.path.join(HERE, "Gd162_GT_open_6sh")
    dst = os.path.join(HERE, "Gd162_finiteT_6sh")
    os.makedirs(dst, exist_ok=True)
    wd = tempfile.mkdtemp()
    s = open(os.path.join(src, "hfbtho_NAMELIST.dat")).read()
    for a, b in (("restart_file = 1", "restart_file = -1"), ("set_temperature = .false.", "set_temperature = .true."),
                 ("temperature = 0.0", "temperature = 0.8")):
        assert a in s
        s = s.replace(a, b)
    open(os.path.join(wd, "hfbtho_NAMELIST.dat"), "w").write(s)
    shutil.copy(os.path.join(src, "hfbtho_output.hel"), wd)
    os.chmod(os.path.join(wd, "hfbtho_output.hel"), 0o644)
    out, t = refrun.run_hfbtho(wd, threads=4)
    assert "iteration converged" in out, out[-3000:]
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)
    points = {}
    for op, k, w, mi in POINTS:
        name = "%s-K%d" % (op, k)
        nml = FAM.format(name=name, re=repr(w.real), im=repr(w.imag), op=op, k=k, max_iter=mi)
        open(os.path.join(wd, name + ".in"), "w").write(nml)
        dat, wall, o = refrun.run_pnfam(wd, name + ".in", threads=4)
        assert "Strength" in dat["rows"], o[-2000:]
        if not points:
            open(os.path.join(dst, "reference_stdout.txt"), "w").write(o)
        points.setdefault(name, []).append({
            "point": "%06d" % len(points.get(name, [])), "namelist": nml,
            "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
            "iters": dat["iters"], "conv": dat["conv"],
            "trace": [[t_[0], t_[1], t_[2], t_[3], t_[4]] for t_ in dat["trace"]], "header": dat["header"]})
        print(name, w, dat["rows"]["Strength"], dat["iters"], flush=True)
    json.dump({"source": "generated with the reference's prebuilt hfbtho_main / pnfam_main.x by tests/golden/make_finite_temp.py",
               "points": points}, open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
