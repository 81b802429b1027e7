#!/usr/bin/env python
"""Known answers for the full-FAM two-body-current field GENERATOR (effective_2bc_extfield, pnfam_extfield_2bc.f90:26-465):
the reference's own prebuilt pnfam_main.x (oracle/_ref) is started in a run directory WITHOUT a .tbc file, so it computes
the Yukawa part of the Gamow-Teller field from scratch and caches it in <name>.tbc (mode 111100).  Cases the reference's
golden trees do not hold: the momentum-dependent terms (two_body_current_usep), beta+, a deformed nucleus.
-> tests/golden/tbc_generator/<case>/{hfbtho_NAMELIST.dat, hfbtho_output.hel, <name>.in, <name>.tbc} + strengths.json
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

CASES = [  # (case, source tree, operator K, beta, usep, omega)
    ("S40_usep_K0", "S40_All_GT2bc", 0, "-", True, 2.0 + 1.0j),
    ("S40_usep_K1", "S40_All_GT2bc", 1, "-", True, 2.0 + 1.0j),
    ("S40_betaplus_K1", "S40_All_GT2bc", 1, "+", False, 3.0 + 0.5j),
    ("S40_betaplus_usep_K0", "S40_All_GT2bc", 0, "+", True, 3.0 + 0.5j),
    ("Gd162_6sh_usep_K1", "Gd162_GT_open_6sh", 1, "-", True, 1.5 + 0.75j),
    ("Gd162_6sh_K0", "Gd162_GT_open_6sh", 0, "-", False, 1.5 + 0.75j),
    ("Gd163_blocked_K0", "Gd163_blocked_6sh", 0, "-", False, 1.5 + 0.75j),      # odd-A, equal-filling blocking
    ("Gd163_blocked_usep_K0", "Gd163_blocked_6sh", 0, "-", True, 1.5 + 0.75j),
    ("Gd162_finiteT_K1", "Gd162_finiteT_6sh", 1, "-", False, 1.5 + 0.75j),      # T = 0.8 MeV
    ("S40_Kminus1", "S40_All_GT2bc", -1, "-", False, 2.0 + 1.0j),                # K = -1 branch of the spatial components
    ("Gd162_6sh_usep_Kminus1", "Gd162_GT_open_6sh", -1, "+", True, 1.5 + 0.75j),
    # a basis size of BASELINE.json configs[4]: about an hour of the reference on 6 threads (the .tbc is 0.9 MB)
    ("Gd162_12sh_K0", "Gd162_SKOP_12sh", 0, "-", False, 2.0 + 1.0j),
    ("Gd162_12sh_K1", "Gd162_SKOP_12sh", 1, "-", False, 2.0 + 1.0j),
    ("Gd162_12sh_usep_K0", "Gd162_SKOP_12sh", 0, "-", True, 2.0 + 1.0j),
]


def main():
    out = os.path.join(HERE, "tbc_generator")
    os.makedirs(out, exist_ok=True)
    spath = os.path.join(out, "strengths.json")
    strengths = json.load(open(spath))["cases"] if os.path.isfile(spath) else {}
    for case, tree, k, beta, usep, w in CASES:
        if len(sys.argv) > 1 and case not in sys.argv[1:]:      # usage: make_tbc_generator.py [case ...]
            continue
        dst = os.path.join(out, case)
        os.makedirs(dst, exist_ok=True)
        tmp = tempfile.mkdtemp()
        for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
            shutil.copy(os.path.join(HERE, tree, f), tmp)
            shutil.copy(os.path.join(HERE, tree, f), dst)
        wd = tempfile.mkdtemp()
        refrun.stage(wd, tmp)
        name = "GT-K%d" % k if k >= 0 else "GT-Km%d" % -k
        nml = FAM.format(name=name, re=repr(w.real), im=repr(w.imag), op="GT", k=k, max_iter=300)
        nml = nml.replace("two_body_current_mode = 0", "two_body_current_mode = 111100")
        nml = nml.replace("beta_type = '-'", "beta_type = '%s'" % beta)
        if usep:
            nml = nml.replace("two_body_current_usep = .false.", "two_body_current_usep = .true.")
        assert "beta_type = '%s'" % beta in nml
        open(os.path.join(wd, name + ".in"), "w").write(nml)
        assert not os.path.exists(os.path.join(wd, name + ".tbc"))
        dat, wall, log = refrun.run_pnfam(wd, name + ".in", threads=int(os.environ.get("REF_THREADS", "4")), timeout=4 * 3600)
        assert "Calculating 2BC matrix elements" in log, log[-2000:]
        shutil.copy(os.path.join(wd, name + ".tbc"), dst)
        open(os.path.join(dst, name + ".in"), "w").write(nml)
        strengths[case] = {"name": name, "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
                           "iters": dat["iters"], "conv": dat["conv"], "wall_s": wall}
        print(case, dat["rows"]["Strength"], dat["iters"], "%.1f s" % wall, flush=True)
        # merge into the file as it is NOW (several of these scripts may run side by side)
        merged = json.load(open(spath))["cases"] if os.path.isfile(spath) else {}
        merged[case] = strengths[case]
        json.dump({"source": "reference's prebuilt pnfam_main.x (oracle/_ref), started without a .tbc file, by "
                             "tests/golden/make_tbc_generator.py", "cases": merged}, open(spath, "w"), indent=0)


if __name__ == "__main__":
    main()
