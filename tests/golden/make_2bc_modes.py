#!/usr/bin/env python
"""Known answers for the closed-form two-body-current modes (nuclear matter + LDA; pnfam_extfield.f90:157-200,
349-366, 562-607, 976-1410), produced with the reference's own prebuilt pnfam_main.x (oracle/_ref) on the 6-shell 40S
case of tests/golden/S40_GT_All: GT with the symmetric / asymmetric nuclear-matter exchange term (2nd digit 2, 3), RS0 /
RS1 / RS2 with the (1 - correction) weight (4th digit 2, 3), P and PS0 with their currents (5th / 6th digit 1),
1BC+2BC and 2BC only, cross-terms on.  -> tests/golden/S40_2bc_modes/points.json
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

CASES = [  # (operator, K, mode, omega)
    ("GT", 0, 121100, 2.0 + 1.0j), ("GT", 1, 131100, 3.0 - 1.5j), ("GT", 0, 221100, 2.0 + 1.0j),
    ("RS0", 0, 121211, 4.0 + 2.0j), ("RS1", 1, 121311, 5.0 + 1.0j), ("RS2", 2, 121200, 3.0 + 2.0j),
    ("P", 1, 121110, 6.0 + 1.5j), ("P", 0, 221110, 6.0 + 1.5j), ("PS0", 0, 121101, 4.0 - 2.0j),
    ("R", 0, 121211, 5.0 + 2.5j), ("RS1", 0, 221311, 2.5 + 1.0j),
]


def main():
    src = os.path.join(HERE, "S40_GT_All")
    dst = os.path.join(HERE, "S40_2bc_modes")
    os.makedirs(dst, exist_ok=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(src, f), dst)
    points = {}
    for op, k, mode, w in CASES:
        wd = tempfile.mkdtemp()
        refrun.stage(wd, dst)
        name = "%s-K%d" % (op, k)
        nml = FAM.format(name=name, re=repr(w.real), im=repr(w.imag), op=op, k=k, max_iter=300)
        nml = nml.replace("two_body_current_mode = 0", "two_body_current_mode = %d" % mode)
        open(os.path.join(wd, name + ".in"), "w").write(nml)
        dat, wall, out = refrun.run_pnfam(wd, name + ".in", threads=1)
        assert "Strength" in dat["rows"], out[-1500:]
        key = "%s-%d" % (name, mode)
        points[key] = [{"point": "000000", "namelist": nml, "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
                        "iters": dat["iters"], "conv": dat["conv"],
                        "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]], "header": dat["header"]}]
        print(key, dat["rows"]["Strength"], dat["iters"], dat["conv"], flush=True)
    json.dump({"source": "reference's prebuilt pnfam_main.x (oracle/_ref) by tests/golden/make_2bc_modes.py", "points": points},
              open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
