#!/usr/bin/env python
"""Generate tests/golden/Gd162_SKOP_20sh/: a 20-shell, 40x40-grid deformed 162Gd case.

Blocks at 20 shells have spin segments longer than one 48-row chunk, which exercises the multi-chunk
accumulation paths of the density / projection kernels.  The reference ships no fixture above 6-8 HO shells
(SURVEY.md section 4), so this one is produced with
the reference's OWN prebuilt executables (oracle/_ref, see oracle/Makefile), in the build container:
  1. hfbtho_main   : HFB ground state, namelist = the reference's
                     tests/"Gd162 closed tests"/GT/000000/hfb_soln/hfbtho_NAMELIST.dat with
                     number_of_shells=20, number_gauss=number_laguerre=40, number_legendre=80,
                     prolate start (restart_file=2, beta2=0.3, basis_deformation=0.3)
  2. pnfam_main.x  : known answers for a few (operator, omega, max_iter) points.
Outputs: hfbtho_NAMELIST.dat, hfbtho_output.hel, points.json (same schema as make_golden.py).
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

SRC = "/root/reference/tests/Gd162 closed tests/GT/000000/hfb_soln/hfbtho_NAMELIST.dat"
FAM = """&general
    fam_output_filename = '{name}'
    print_stdout = .true.
    use_fam_storage = 0
    real_eqrpa = {re}
    imag_eqrpa = {im}
/

&ext_field
    beta_type = '-'
    operator_name = '{op}'
    operator_k = {k}
    compute_crossterms = .true.
    two_body_current_mode = 0
    two_body_current_usep = .false.
    two_body_current_lecs = -3.1962, 3.1962, 0
/

&interaction
    interaction_name = 'SKOP'
    require_self_consistency = .true.
    require_gauge_invariance = .true.
    force_j2_terms = .false.
    vpair_t0 = -346.352
    vpair_t1 = ,
    override_cs0 = 128.279
    override_csr = 0.0
    override_cds = 0.0
    override_ct = ,
    override_cgs = ,
    override_cf = ,
/

&solver
    max_iter = {max_iter}
    convergence_epsilon = 1e-07
    broyden_history_size = 50
    energy_shift_prot = 0.0
    energy_shift_neut = 0.0
    quench_residual_int = 1.0
/
"""
POINTS = [  # (operator, K, omega, max_iter) -- few iterations: one reference iteration takes ~5 s at 20 shells
    ("GT", 0, 2.0 + 1.0j, 3),
    ("GT", 1, 5.0 + 0.5j, 3),
    ("RS1", 1, 3.0 + 2.0j, 2),
]


def main():
    dst = os.path.join(HERE, "Gd162_SKOP_20sh")
    os.makedirs(dst, exist_ok=True)
    wd = tempfile.mkdtemp()
    if not os.path.isfile(os.path.join(dst, "hfbtho_output.hel")):
        s = open(SRC).read()
        for a, b in (("number_of_shells = 6", "number_of_shells = 20"), ("number_gauss = 20", "number_gauss = 40"),
                     ("number_laguerre = 20", "number_laguerre = 40"), ("number_legendre = 40", "number_legendre = 80"),
                     ("restart_file = 1", "restart_file = 2"), ("beta2_deformation = 0.0", "beta2_deformation = 0.3"),
                     ("basis_deformation = 0.0", "basis_deformation = 0.3")):
            assert a in s
            s = s.replace(a, b)
        open(os.path.join(wd, "hfbtho_NAMELIST.dat"), "w").write(s)
        out, t = refrun.run_hfbtho(wd, threads=os.cpu_count())
        assert "iteration converged" in out, out[-2000:]
        for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
            shutil.copy(os.path.join(wd, f), dst)
    else:
        for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
            shutil.copy(os.path.join(dst, f), wd)
    points = {}
    for op, k, w, mi in POINTS:
        name = "%s-K%d" % (op, k)
        nml = FAM.format(name=name, re=repr(w.real), im=repr(w.imag), op=op, k=k, max_iter=mi)
        open(os.path.join(wd, name + ".in"), "w").write(nml)
        dat, wall, out = refrun.run_pnfam(wd, name + ".in", threads=os.cpu_count())
        assert "Strength" in dat["rows"], out[-2000:]
        points.setdefault(name, []).append({
            "point": "%06d" % len(points.get(name, [])), "namelist": nml,
            "rows": {kk: [repr(v.real), repr(v.imag)] for kk, v in dat["rows"].items()},
            "iters": dat["iters"], "conv": dat["conv"],
            "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]], "header": dat["header"],
            "ref_wall_s": wall, "ref_threads": os.cpu_count(),
        })
        print(name, w, mi, dat["rows"]["Strength"], dat["iters"], "wall %.1fs" % wall, flush=True)
    json.dump({"source": "generated with the reference's prebuilt hfbtho_main / pnfam_main.x by tests/golden/make_gd162_16sh.py",
               "points": points}, open(os.path.join(dst, "points.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
