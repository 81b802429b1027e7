"""CPU tests of the full-FAM two-body-current field generator (csrc/host/tbc_generator.cpp; the reference's
effective_2bc_extfield, exes/pnfam/pnfam_extfield_2bc.f90:26-465) -- SURVEY.md section 8f row 3.

Known answers: the <OP>.tbc files the reference itself wrote,
  * tests/golden/S40_All_GT2bc/GT-K{0,1}.tbc: from the reference's own golden tree tests/S40_All_GT2bc/hfb_soln,
  * tests/golden/tbc_generator/*: the reference's prebuilt pnfam_main.x started here without a .tbc file
    (tests/golden/make_tbc_generator.py): momentum-dependent terms, beta+, a deformed nucleus, a blocked odd-A nucleus,
    finite temperature, K = -1, and 162Gd at 12 shells.
A run directory WITHOUT the .tbc file makes the library compute the field and cache it in the reference's record layout;
the file is compared record by record, the field element by element, and the strengths of the complete FAM solve
(CPU oracle) with the reference's.  Bound: 1e-12 of the largest element of a component (measured: 7e-15)."""
import json
import os
import re
import shutil
import struct

import numpy as np
import pytest

from conftest import GOLDEN, gold_rows, load_points
from oracle import fam_oracle as fo
from pynfam_b200 import host

GEN = os.path.join(GOLDEN, "tbc_generator")
GEN_CASES = ["S40_usep_K0", "S40_usep_K1", "S40_betaplus_K1", "S40_betaplus_usep_K0", "Gd162_6sh_usep_K1", "Gd162_6sh_K0",
             "Gd163_blocked_K0", "Gd163_blocked_usep_K0", "Gd162_finiteT_K1", "S40_Kminus1", "Gd162_6sh_usep_Kminus1",
             "Gd162_12sh_K0", "Gd162_12sh_K1", "Gd162_12sh_usep_K0"]    # 12 shells: a basis size of BASELINE.json configs[4]; 9-22 minutes of the reference each


def records(path):
    d = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(d):
        n = struct.unpack("<i", d[pos:pos + 4])[0]
        out.append(d[pos + 4:pos + 4 + n])
        assert struct.unpack("<i", d[pos + 4 + n:pos + 8 + n])[0] == n
        pos += n + 8
    return out


def compare_files(mine, ref):
    a, b = records(mine), records(ref)
    assert [len(r) for r in a] == [len(r) for r in b]
    worst = 0.0
    for i, (x, y) in enumerate(zip(a, b)):
        if len(x) < 1000:
            assert x == y, ("header record", i)        # version, key, flags, LECs, dimensions, label, K, ...
            continue
        x, y = np.frombuffer(x, "<f8"), np.frombuffer(y, "<f8")
        scale = np.abs(y).max()
        if scale < 1e-14:                              # the c4 direct and momentum direct parts vanish identically
            assert np.abs(x).max() < 1e-14
            continue
        worst = max(worst, np.abs(x - y).max() / scale)
    return worst


def stage(src_dir, wd, namelist_text, name):
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(src_dir, f), wd)
    with open(os.path.join(wd, name + ".in"), "w") as f:
        f.write(namelist_text)


@pytest.mark.parametrize("op", ["GT-K0", "GT-K1"])
def test_generated_file_equals_the_golden_tree_file(op, tmp_path):
    wd = str(tmp_path)
    src = os.path.join(GOLDEN, "S40_All_GT2bc")
    stage(src, wd, load_points("S40_All_GT2bc")[op][3]["namelist"], op)
    p = host.Problem(wd, op + ".in")
    assert os.path.isfile(os.path.join(wd, op + ".tbc"))
    worst = compare_files(os.path.join(wd, op + ".tbc"), os.path.join(src, op + ".tbc"))
    print("worst relative difference of a component: %.2e" % worst)
    assert worst < 1e-12
    # the field set up from the generated elements equals the one set up from the reference's file
    wd2 = str(tmp_path / "from_file")
    os.makedirs(wd2)
    stage(src, wd2, load_points("S40_All_GT2bc")[op][3]["namelist"], op)
    shutil.copy(os.path.join(src, op + ".tbc"), wd2)
    q = host.Problem(wd2, op + ".in")
    f1, f2 = p.f64("f_elem"), q.f64("f_elem")
    assert np.abs(f1 - f2).max() < 1e-12 * np.abs(f2).max()


@pytest.mark.parametrize("case", GEN_CASES)
def test_generated_file_equals_the_reference_binary_file(case, tmp_path):
    wd = str(tmp_path)
    src = os.path.join(GEN, case)
    info = json.load(open(os.path.join(GEN, "strengths.json")))["cases"][case]
    name = info["name"]
    stage(src, wd, open(os.path.join(src, name + ".in")).read(), name)
    host.Problem(wd, name + ".in")
    worst = compare_files(os.path.join(wd, name + ".tbc"), os.path.join(src, name + ".tbc"))
    print(case, "worst relative difference of a component: %.2e" % worst)
    assert worst < 1e-12


@pytest.mark.parametrize("case", ["S40_usep_K1", "Gd162_6sh_K0", "Gd163_blocked_K0"])
def test_plain_array_entry_point_equals_the_reference_file(case, tmp_path):
    """pnfam_host_effective_2bc_extfield (include/pnfam_b200.h): the entry a Fortran maintainer calls in place of
    effective_2bc_extfield, fed with the arrays that routine takes from its modules (HFBTHO quantum numbers, oscillator
    lengths, rk, the block structure of the operator) -- against the .tbc the reference binary wrote."""
    wd = str(tmp_path)
    src = os.path.join(GEN, case)
    info = json.load(open(os.path.join(GEN, "strengths.json")))["cases"][case]
    name = info["name"]
    nml = open(os.path.join(src, name + ".in")).read()
    stage(src, wd, nml, name)
    shutil.copy(os.path.join(src, name + ".tbc"), wd)          # the set-up itself reads the reference's file: no generation
    os.environ["PNFAM_B200_NO_TBC_GENERATOR"] = "1"
    try:
        p = host.Problem(wd, name + ".in")
    finally:
        del os.environ["PNFAM_B200_NO_TBC_GENERATOR"]
    nb = p.iscalar("hfb_nb")
    rk = p.f64("hfb_rk").reshape(2 * nb, -1).T                  # Fortran (nqx, 2 nbx)
    out = host.effective_2bc_extfield(p.i32("hfb_id"), p.i32("hfb_nz"), p.i32("hfb_nr"), p.i32("hfb_nl"), p.i32("hfb_ns"),
                                      p.scalar("hfb_bz"), p.scalar("hfb_bp"), rk, p.i32("f_ir2c"), p.i32("f_ir2m"), p.iscalar("nxy"),
                                      k=int(re.search(r"operator_k\s*=\s*(-?\d+)", nml).group(1)), beta_minus=bool(p.iscalar("beta_minus")),
                                      use_p=".true." in nml.split("two_body_current_usep")[1].split("\n")[0])
    ref = [np.frombuffer(r, "<f8") for r in records(os.path.join(src, name + ".tbc")) if len(r) > 1000]
    for a, b in zip(out, ref):
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= 1e-12 * max(scale, 1e-2)
    # unsorted variant: the same numbers in the original in-block order (a permutation of every block)
    raw = host.effective_2bc_extfield(p.i32("hfb_id"), p.i32("hfb_nz"), p.i32("hfb_nr"), p.i32("hfb_nl"), p.i32("hfb_ns"),
                                      p.scalar("hfb_bz"), p.scalar("hfb_bp"), rk, p.i32("f_ir2c"), p.i32("f_ir2m"), p.iscalar("nxy"),
                                      k=int(re.search(r"operator_k\s*=\s*(-?\d+)", nml).group(1)), beta_minus=bool(p.iscalar("beta_minus")),
                                      use_p=False, spin_sorted=False)
    assert abs(np.sort(np.abs(raw[0])) - np.sort(np.abs(out[0]))).max() < 1e-14
    with pytest.raises(host.PnfamError, match="invalid K"):
        host.effective_2bc_extfield(p.i32("hfb_id"), p.i32("hfb_nz"), p.i32("hfb_nr"), p.i32("hfb_nl"), p.i32("hfb_ns"),
                                    p.scalar("hfb_bz"), p.scalar("hfb_bp"), rk, p.i32("f_ir2c"), p.i32("f_ir2m"), p.iscalar("nxy"), k=2)


def test_second_set_up_reads_the_cached_file(tmp_path):
    """The generated file is what the next process reads (as the reference does): same field, no recomputation."""
    wd = str(tmp_path)
    src = os.path.join(GOLDEN, "S40_All_GT2bc")
    nml = load_points("S40_All_GT2bc")["GT-K1"][0]["namelist"]
    stage(src, wd, nml, "GT-K1")
    f1 = host.Problem(wd, "GT-K1.in").f64("f_elem")
    stamp = os.path.getmtime(os.path.join(wd, "GT-K1.tbc"))
    os.environ["PNFAM_B200_NO_TBC_GENERATOR"] = "1"      # a second set-up must not need the generator
    try:
        f2 = host.Problem(wd, "GT-K1.in").f64("f_elem")
    finally:
        del os.environ["PNFAM_B200_NO_TBC_GENERATOR"]
    assert os.path.getmtime(os.path.join(wd, "GT-K1.tbc")) == stamp
    assert np.array_equal(f1, f2)


@pytest.mark.parametrize("op,idx", [("GT-K0", 4), ("GT-K1", 12)])
def test_fam_solve_with_generated_field_matches_golden_point(op, idx, tmp_path):
    """Complete chain without any reference-made file: generated field -> FAM solve (CPU oracle) -> the golden strength,
    cross-terms and iteration count of tests/S40_All_GT2bc (1e-9)."""
    wd = str(tmp_path)
    pt = load_points("S40_All_GT2bc")[op][idx]
    stage(os.path.join(GOLDEN, "S40_All_GT2bc"), wd, pt["namelist"], "x")
    p = host.Problem(wd, "x.in")
    it, si, st = fo.solver_from_problem(p).solve(p.iscalar("max_iter"), p.scalar("convergence_epsilon"))
    assert it == pt["iters"]
    gold = gold_rows(pt)
    labels = ["Strength"] + [p.label(i) for i in range(1, 1 + p.iscalar("nxterms"))]
    for k, lab in enumerate(labels):
        if lab in gold:
            assert abs(st[k] - gold[lab]) <= 1e-9 * abs(gold[lab]), (lab, st[k], gold[lab])


@pytest.mark.parametrize("case", ["S40_usep_K0", "S40_betaplus_K1", "Gd162_6sh_usep_K1", "Gd163_blocked_usep_K0", "Gd162_finiteT_K1"])
def test_fam_solve_with_generated_field_matches_reference_binary(case, tmp_path):
    """The same for the momentum-dependent terms, beta+, the deformed nucleus, the blocked odd-A nucleus (density matrix with
    the equal-filling term) and finite temperature (thermal density matrix): strengths of the reference binary that
    computed its own field (tests/golden/tbc_generator/strengths.json)."""
    wd = str(tmp_path)
    src = os.path.join(GEN, case)
    info = json.load(open(os.path.join(GEN, "strengths.json")))["cases"][case]
    name = info["name"]
    stage(src, wd, open(os.path.join(src, name + ".in")).read(), name)
    p = host.Problem(wd, name + ".in")
    it, si, st = fo.solver_from_problem(p).solve(p.iscalar("max_iter"), p.scalar("convergence_epsilon"))
    assert it == info["iters"]
    labels = ["Strength"] + [p.label(i) for i in range(1, 1 + p.iscalar("nxterms"))]
    for k, lab in enumerate(labels):
        g = complex(float(info["rows"][lab][0]), float(info["rows"][lab][1]))
        assert abs(st[k] - g) <= 1e-9 * abs(g), (case, lab, st[k], g)


def test_generator_can_be_switched_off(tmp_path):
    wd = str(tmp_path)
    stage(os.path.join(GOLDEN, "S40_All_GT2bc"), wd, load_points("S40_All_GT2bc")["GT-K0"][0]["namelist"], "x")
    os.environ["PNFAM_B200_NO_TBC_GENERATOR"] = "1"
    try:
        with pytest.raises(host.PnfamError, match="tbc"):
            host.Problem(wd, "x.in")
    finally:
        del os.environ["PNFAM_B200_NO_TBC_GENERATOR"]


def test_factorised_radial_elements_equal_the_literal_sum_at_12_shells(tmp_path):
    """Beyond the sizes the reference binary can be run at here: the per-(A, C) intermediate of the radial elements against
    the reference's literal four-fold Cartesian sum (PNFAM_B200_TBC_LITERAL_RADIAL=1), 12-shell 162Gd, K = 0; the two
    .tbc files agree to 1e-12 of the largest element."""
    src = os.path.join(GOLDEN, "Gd162_SKOP_12sh")
    info = json.load(open(os.path.join(GEN, "strengths.json")))["cases"]["Gd162_6sh_K0"]
    nml = open(os.path.join(GEN, "Gd162_6sh_K0", info["name"] + ".in")).read()
    files = []
    for sub, env in (("fast", None), ("literal", "1")):
        wd = str(tmp_path / sub)
        os.makedirs(wd)
        stage(src, wd, nml, info["name"])
        if env:
            os.environ["PNFAM_B200_TBC_LITERAL_RADIAL"] = env
        try:
            host.Problem(wd, info["name"] + ".in")
        finally:
            os.environ.pop("PNFAM_B200_TBC_LITERAL_RADIAL", None)
        files.append(os.path.join(wd, info["name"] + ".tbc"))
    worst = compare_files(files[0], files[1])
    print("12 shells, factorised vs literal radial sum: %.2e" % worst)
    assert worst < 1e-12


def test_chunked_z_contraction_gives_the_same_field(tmp_path):
    """PNFAM_B200_TBC_MEM_GB bounds the z-contracted densities kept at once (large bases: 24 shells needs two chunks under
    the default 4 GB); with a bound that forces several chunks at 6 shells the file still equals the reference's."""
    wd = str(tmp_path)
    src = os.path.join(GOLDEN, "S40_All_GT2bc")
    stage(src, wd, load_points("S40_All_GT2bc")["GT-K1"][0]["namelist"], "GT-K1")
    os.environ["PNFAM_B200_TBC_MEM_GB"] = "0.003"
    os.environ["PNFAM_B200_SETUP_TIMING"] = "1"
    fd_saved = os.dup(2)
    log = str(tmp_path / "stderr.txt")
    fd = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)
    try:
        host.Problem(wd, "GT-K1.in")
    finally:
        os.dup2(fd_saved, 2)
        os.close(fd)
        del os.environ["PNFAM_B200_TBC_MEM_GB"], os.environ["PNFAM_B200_SETUP_TIMING"]
    assert open(log).read().count("2BC z contraction") >= 2          # really chunked
    assert compare_files(os.path.join(wd, "GT-K1.tbc"), os.path.join(src, "GT-K1.tbc")) < 1e-12
