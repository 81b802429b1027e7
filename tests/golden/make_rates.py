#!/usr/bin/env python
"""Golden data for the beta-decay rate chain (phase space -> shape factor -> rates) from the reference's own outputs
for 40S: tests/S40_GT_All/000000 (2024 tree, current cross-term definitions, gA = -1.27) and for 162Gd, Gamow-Teller only:
tests/'Gd162 closed tests'/GT/000000.
  fam_soln/*.out.ctr                      the 14 strength files (inputs of shapeFactor)          -> S40_GT_All/fam_soln/
  beta_soln/beta.out                      every rate / half-life row
  beta_soln/beta_meta/phasespace_{re,im}  f1..f6 at the 60 contour points
  beta_soln/shapefactor_im, beta_meta/shapefactor_re   every shape-factor column                -> S40_GT_All/beta_soln.json
  beta_soln/logfile_mini.dat              HFB_Qval, EQRPA_max (the inputs pynfam took from the HFB log)
Data files of the reference's test tree, not source.  Run ONCE in the build container; the results are committed."""
import json
import os
import shutil

TREES = [
    # (reference tree, golden case, operators, settings)
    ("/root/reference/tests/S40_GT_All/000000", "S40_GT_All",
     ["F-K0", "GT-K0", "GT-K1", "RS0-K0", "PS0-K0", "R-K0", "R-K1", "P-K0", "P-K1", "RS1-K0", "RS1-K1", "RS2-K0", "RS2-K1", "RS2-K2"]),
    # heavy deformed nucleus (Z = 64: strong Coulomb distortion), Gamow-Teller only
    ("/root/reference/tests/Gd162 closed tests/GT/000000", "Gd162_GT_closed_6sh", ["GT-K0", "GT-K1"]),
]
HERE = os.path.dirname(os.path.abspath(__file__))


def table(path):
    lines = [ln for ln in open(path).read().split("\n") if ln.strip() and not ln.startswith("#")]
    head = lines[0].split()
    rows = [ln.split()[1:] for ln in lines[1:]]
    return {h: [r[i] for r in rows] for i, h in enumerate(head)}


for SRC, case, OPS in TREES:
    DST = os.path.join(HERE, case)
    os.makedirs(os.path.join(DST, "fam_soln"), exist_ok=True)
    for op in OPS:
        shutil.copy(os.path.join(SRC, "fam_soln", op + ".out.ctr"), os.path.join(DST, "fam_soln", op + ".out.ctr"))
        os.chmod(os.path.join(DST, "fam_soln", op + ".out.ctr"), 0o644)
    rates = {}
    for ln in open(os.path.join(SRC, "beta_soln", "beta.out")):
        t = ln.split()
        if len(t) == 3 and not ln.startswith("#") and t[0] != "Rate(s^-1)":
            rates[t[0]] = {"rate": t[1], "halflife": t[2]}
    log = open(os.path.join(SRC, "beta_soln", "logfile_mini.dat")).read().split("\n")
    hd, row = log[0].split(), log[1].split()
    hfb = {"HFB_Qval": row[1 + hd.index("HFB_Qval")], "EQRPA_max": row[1 + hd.index("EQRPA_max")], "E_gs": row[1 + hd.index("E_gs")]}
    ps_re, ps_im = table(os.path.join(SRC, "beta_soln", "beta_meta", "phasespace_re.out")), table(os.path.join(SRC, "beta_soln", "beta_meta", "phasespace_im.out"))
    sf_re, sf_im = table(os.path.join(SRC, "beta_soln", "beta_meta", "shapefactor_re.out")), table(os.path.join(SRC, "beta_soln", "shapefactor_im.out"))
    out = {"source": "mld1812/pynfam " + SRC[len("/root/reference/"):] + "/beta_soln", "settings": {"GA": -1.27, "GV": 1.0, "psi_glpts": 15, "ratint_pts": 20, "energy_max": float(hfb["EQRPA_max"])},
           "hfb": hfb, "rates": rates,
           "phase_space": {k: {"re": ps_re[k], "im": ps_im[k]} for k in ("f1", "f2", "f3", "f4", "f5", "f6")},
           "shape_factor": {k: {"re": sf_re[k], "im": sf_im[k]} for k in sf_re if k not in ("Re(EQRPA)", "Im(EQRPA)")}}
    json.dump(out, open(os.path.join(DST, "beta_soln.json"), "w"), indent=0)
    print(case, sorted(os.listdir(os.path.join(DST, "fam_soln"))), os.path.getsize(os.path.join(DST, "beta_soln.json")), hfb, len(rates))
