#!/usr/bin/env python
"""Generate tests/golden/S40_custom_interaction/: a FAM run whose coupling constants come from a file
(interaction_name = 'FILE:custom_edf.dat', namelist &CUSTOM_INTERACTION, pnfam_interaction.f90:303-331).  The reference
ships only the empty template exes/pnfam/pnfam_CUSTOM_EDF.dat.  The file written here carries the SkO' couplings of the
S40 fixture (taken at full precision from our own set-up, which is pinned to the reference's header) with a modified
time-odd sector (Cs0, Csr, sigma_s, Cgs) and T=0 pairing, so that a run that ignored the file would not reproduce the
answers.  Known answers: the reference's prebuilt pnfam_main.x (oracle/_ref)."""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import refrun  # noqa: E402
from conftest import stage_point  # noqa: E402
from pynfam_b200 import host  # noqa: E402

POINTS = [("GT", 0, 2.0 + 1.0j), ("GT", 1, 5.0 + 0.5j), ("RS1", 0, 6.0 + 2.0j), ("F", 0, 1.0 + 0.75j)]


def main():
    dst = os.path.join(HERE, "S40_custom_interaction")
    os.makedirs(dst, exist_ok=True)
    wd = tempfile.mkdtemp()
    pt = stage_point("S40_SKOP_6sh", "GT-K0", 10, wd)
    p = host.Problem(wd, "x.in")
    c = {k: p.scalar(k) for k in ("cr0", "crr", "sigma_r", "cdrho", "ctau", "ctj0", "ctj1", "ctj2", "crdj", "cs0", "csr", "sigma_s", "cds", "ct", "cj",
                                   "csdj", "cf", "cgs", "cpair0", "cpairr", "cspair0", "cspairr", "sigma_pair")}
    c.update({"cs0": 110.0, "csr": 25.0, "sigma_s": 0.5, "cgs": 12.5, "cspair0": -20.0})
    c["cspairr"] = c["cspair0"] * (c["cpairr"] / c["cpair0"])      # the reference checks the density dependence of both pairing channels
    order = ["cr0", "crr", "sigma_r", "cdrho", "ctau", "ctj0", "ctj1", "ctj2", "crdj", "cs0", "csr", "sigma_s", "cds", "ct", "cj", "csdj", "cf", "cgs",
             "cpair0", "cpairr", "cspair0", "cspairr", "sigma_pair"]
    with open(os.path.join(wd, "custom_edf.dat"), "w") as f:
        f.write("&CUSTOM_INTERACTION\n" + "".join("   %-10s = %r\n" % (k, float(c[k])) for k in order) + "/\n")
    shutil.copy(os.path.join(wd, "custom_edf.dat"), dst)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(wd, f), dst)
    import re
    base = re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", pt["namelist"])
    points = {}
    for op, k, w in POINTS:
        name = "%s-K%d" % (op, k)
        import re
        nml = base.replace("interaction_name = 'SKOP'", "interaction_name = 'FILE:custom_edf.dat'")
        nml = re.sub(r"operator_name = '\w+'", "operator_name = '%s'" % op, nml)
        nml = re.sub(r"operator_k = \d+", "operator_k = %d" % k, nml)
        nml = re.sub(r"real_eqrpa = \S+", "real_eqrpa = %r" % w.real, nml)
        nml = re.sub(r"imag_eqrpa = \S+", "imag_eqrpa = %r" % w.imag, nml)
        nml = re.sub(r"fam_output_filename = '[^']*'", "fam_output_filename = '%s'" % name, nml)
        assert "FILE:custom_edf.dat" in nml
        open(os.path.join(wd, name + ".in"), "w").write(nml)
        dat, wall, o = refrun.run_pnfam(wd, name + ".in", threads=4)
        assert "Strength" in dat["rows"], o[-3000:]
        points.setdefault(name, []).append({
            "point": "000000", "namelist": nml, "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
            "iters": dat["iters"], "conv": dat["conv"], "trace": [[t_[0], t_[1], t_[2], t_[3], t_[4]] for t_ in dat["trace"]], "header": dat["header"]})
        print(name, w, dat["rows"]["Strength"], dat["iters"], flush=True)
    json.dump({"source": "generated with the reference's prebuilt pnfam_main.x by tests/golden/make_custom_interaction.py", "points": points},
              open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
