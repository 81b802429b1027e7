#!/usr/bin/env python
"""Known answers for the density-matrix-expansion two-body-current modes (pnfam_extfield.f90:170-181, 349-366, 562-607,
1195-1330, 1380-1662; tau and Delta rho from HFBTHO's DENSIT), produced with the reference's own prebuilt pnfam_main.x
(oracle/_ref): GT with the DME exchange term alone (2nd digit 4) and with the direct part of the full-FAM field on top
(2nd digit 5: the reference computes its .tbc from scratch), P with the DME vector current (5th digit 2), PS0 with the DME
axial charge (6th digit 2); 1BC+2BC and 2BC only, cross-terms on; spherical 40S and deformed 162Gd at 6 shells,
blocked 163Gd and 162Gd at T = 0.8 MeV.
-> tests/golden/S40_2bc_dme/points.json, tests/golden/Gd162_2bc_dme/points.json
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

CASES = {
    "S40_2bc_dme": ("S40_GT_All", [  # (operator, K, mode, omega)
        ("GT", 0, 141100, 2.0 + 1.0j), ("GT", 1, 141100, 3.0 - 1.5j), ("GT", 0, 241100, 2.0 + 1.0j),
        ("GT", 0, 151100, 2.0 + 1.0j), ("GT", 1, 151100, 3.0 - 1.5j),
        ("P", 0, 121120, 6.0 + 1.5j), ("P", 1, 121120, 6.0 + 1.5j), ("P", 0, 221120, 6.0 + 1.5j),
        ("PS0", 0, 121102, 4.0 - 2.0j), ("PS0", 0, 221102, 4.0 - 2.0j), ("RS0", 0, 141222, 4.0 + 2.0j)]),
    "Gd162_2bc_dme": ("Gd162_GT_open_6sh", [
        ("GT", 1, 141100, 1.5 + 0.75j), ("GT", 0, 151100, 1.5 + 0.75j), ("P", 1, 121120, 5.0 + 1.0j), ("PS0", 0, 121102, 4.0 + 1.0j)]),
    # tau and Delta rho with the equal-filling term of the blocked level / the thermal occupations
    "Gd163_2bc_dme": ("Gd163_blocked_6sh", [("GT", 0, 141100, 1.5 + 0.75j), ("GT", 1, 151100, 1.5 + 0.75j), ("PS0", 0, 121102, 4.0 + 1.0j)]),
    "Gd162T_2bc_dme": ("Gd162_finiteT_6sh", [("GT", 1, 141100, 1.5 + 0.75j), ("P", 0, 121120, 5.0 + 1.0j)]),
}


def main():
    for out, (tree, cases) in CASES.items():
        if len(sys.argv) > 1 and out not in sys.argv[1:]:      # usage: make_2bc_dme.py [group ...]
            continue
        dst = os.path.join(HERE, out)
        os.makedirs(dst, exist_ok=True)
        for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
            shutil.copy(os.path.join(HERE, tree, f), dst)
        points = {}
        for op, k, mode, w in cases:
            wd = tempfile.mkdtemp()
            refrun.stage(wd, dst)
            name = "%s-K%d" % (op, k)
            nml = FAM.format(name=name, re=repr(w.real), im=repr(w.imag), op=op, k=k, max_iter=300)
            nml = nml.replace("two_body_current_mode = 0", "two_body_current_mode = %d" % mode)
            open(os.path.join(wd, name + ".in"), "w").write(nml)
            dat, wall, log = refrun.run_pnfam(wd, name + ".in", threads=4)
            assert "Strength" in dat["rows"], log[-1500:]
            key = "%s-%d" % (name, mode)
            points[key] = [{"point": "000000", "namelist": nml, "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
                            "iters": dat["iters"], "conv": dat["conv"],
                            "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]], "header": dat["header"]}]
            print(out, key, dat["rows"]["Strength"], dat["iters"], dat["conv"], "%.1f s" % wall, flush=True)
        json.dump({"source": "reference's prebuilt pnfam_main.x (oracle/_ref) by tests/golden/make_2bc_dme.py", "points": points},
                  open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
