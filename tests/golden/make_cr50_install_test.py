#!/usr/bin/env python
"""The reference's own install regression test (exes/pnfam/tests/pnfam2_serial/test_nompi.sh): 50Cr, SLy4, 6 shells --
HFBTHO ground state (E_HFB = -432.187650 expected to 0.1 %), then the GT- K=1 strength function at omega = (2, 6, 10) + 2i MeV
with the published values 1.0604161621265991E-01, 1.2528722692977165E-01, 4.0138005453632780E-01 (0.1 %).
Here: the reference's prebuilt hfbtho_main (oracle/_ref) makes the HFB solution from the test's own namelist, and its
pnfam_main.x the three points (for the tight comparison).  -> tests/golden/Cr50_SLY4_6sh/{hfbtho_NAMELIST.dat,
hfbtho_output.hel, points.json}
"""
import json
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refrun  # noqa: E402

SRC = "/root/reference/exes/pnfam/tests/pnfam2_serial"
PUBLISHED = {2.0: 1.0604161621265991E-01, 6.0: 1.2528722692977165E-01, 10.0: 4.0138005453632780E-01}   # test_nompi.sh:139-156


def main():
    dst = os.path.join(HERE, "Cr50_SLY4_6sh")
    os.makedirs(dst, exist_ok=True)
    wd = tempfile.mkdtemp()
    shutil.copy(os.path.join(SRC, "hfbtho_NAMELIST.dat"), wd)
    log, wall = refrun.run_hfbtho(wd, threads=4)
    tho = open(os.path.join(wd, "thoout.dat"), errors="replace").read()
    ehfb = float(re.findall(r"tEnergy: ehfb \(qp\)\.\.\.\s+(\S+)", tho)[-1])
    print("E_HFB", ehfb, "(expected -432.187650)", "%.1f s" % wall, flush=True)
    assert abs(ehfb / -432.187650 - 1) < 1e-3
    # the zero-iteration namelist pnFAM needs is the same file (pnFAM forces number_iterations = 0 itself)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)
    base = open(os.path.join(SRC, "pnfam_NAMELIST.dat")).read()
    pts = []
    for i, w in enumerate(sorted(PUBLISHED)):
        run = tempfile.mkdtemp()
        refrun.stage(run, dst)
        nml = base.replace('fam_output_filename = ""', "fam_output_filename = 'GT-K1'").replace("real_eqrpa = 10.0", "real_eqrpa = %r" % w)
        open(os.path.join(run, "GT-K1.in"), "w").write(nml)
        dat, wall, out = refrun.run_pnfam(run, "GT-K1.in", threads=1)
        s = dat["rows"]["Strength"]
        # the test reads word 2 of the OP.out line: the imaginary part of S (columns: energy, Re S, Im S)
        print(w, s, dat["iters"], "published", PUBLISHED[w], "rel", abs(s.imag / PUBLISHED[w] - 1), flush=True)
        assert abs(s.imag / PUBLISHED[w] - 1) < 1e-3
        pts.append({"point": "%06d" % i, "namelist": nml, "rows": {k: [repr(v.real), repr(v.imag)] for k, v in dat["rows"].items()},
                    "iters": dat["iters"], "conv": dat["conv"], "published_im_strength": PUBLISHED[w],
                    "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]], "header": dat["header"]})
    json.dump({"source": "exes/pnfam/tests/pnfam2_serial (reference's install test) run with oracle/_ref by "
                         "tests/golden/make_cr50_install_test.py", "ehfb": ehfb, "points": {"GT-K1": pts}},
              open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
