"""CPU tests of the host logic and of the C ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_points, stage_point
from oracle import refrun
from pynfam_b200 import host


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pnfam_b200.h")).read()
    return sorted(set(re.findall(r"\b(pnfam_(?:b200_|problem_)\w+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported():
    syms = _declared_symbols()
    assert len(syms) >= 12
    libs = [ctypes.CDLL(os.path.join(ROOT, "pynfam_b200", "lib", n)) for n in ("libpnfam_host.so", "libpnfam_b200.so")]
    for s in syms:
        assert any(hasattr(l, s) for l in libs), s


def test_gpu_library_fails_loudly_without_device():
    """No CPU fallback: without a CUDA device the DMMA probe (and every other entry) must return an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pynfam_b200 import gpu
    with pytest.raises(gpu.GpuError, match="no CUDA device|CUDA"):
        gpu.dmma_peak_tflops()


def test_namelist_roundtrip(tmp_path):
    g = {"general": {"fam_output_filename": "GT-K0", "print_stdout": True, "real_eqrpa": 0.5, "imag_eqrpa": -2.25},
         "interaction": {"interaction_name": "SKOP", "vpair_t1": None, "override_cs0": 128.279}}
    path = str(tmp_path / "a.in")
    refrun.write_namelist(path, g)
    back = refrun.read_namelist(path)
    assert back["general"]["fam_output_filename"] == "GT-K0"
    assert back["general"]["imag_eqrpa"] == -2.25
    assert back["interaction"]["vpair_t1"] is None


def test_namelists_with_several_assignments_per_line(tmp_path):
    """The namelists shipped with HFBTHO and the reference's own install test (exes/pnfam/tests/pnfam2_serial) put several
    `key = value` pairs on a line and close the group on the same line; pynfam writes one per line.  Both layouts of the
    same content must set up the same problem (neutron number, shells, functional, pairing ...)."""
    import shutil
    g = os.path.join(GOLDEN, "Cr50_SLY4_6sh")
    pt = load_points("Cr50_SLY4_6sh")["GT-K1"][0]
    text = open(os.path.join(g, "hfbtho_NAMELIST.dat")).read()
    assert "proton_number = 24, neutron_number = 26, type_of_calculation = 1 /" in text      # the multi-key layout
    one, multi = tmp_path / "one", tmp_path / "multi"
    for d in (one, multi):
        d.mkdir()
        shutil.copy(os.path.join(g, "hfbtho_output.hel"), d)
        (d / "x.in").write_text(pt["namelist"])
    (multi / "hfbtho_NAMELIST.dat").write_text(text)
    # one assignment per line, '/' on its own line
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("&"):
            head, _, rest = line.partition(" ")
            out.append(head)
            line = rest.strip()
        closing = line.endswith("/")
        line = line[:-1].strip() if closing else line
        parts, cur, depth = [], "", 0
        for tok in line.split(","):
            if "=" in tok and cur:
                parts.append(cur)
                cur = tok
            else:
                cur = cur + "," + tok if cur else tok
        if cur.strip():
            parts.append(cur)
        out += ["    " + q.strip() for q in parts if q.strip()]
        if closing:
            out.append("/")
    (one / "hfbtho_NAMELIST.dat").write_text("\n".join(out) + "\n")
    a, b = host.Problem(str(one), "x.in"), host.Problem(str(multi), "x.in")
    assert a.iscalar("npr_n") == b.iscalar("npr_n") == 26 and a.iscalar("npr_p") == b.iscalar("npr_p") == 24
    assert a.iscalar("n_shells") == b.iscalar("n_shells") == 6
    assert np.array_equal(a.f64("En"), b.f64("En")) and np.array_equal(a.f64("f_elem"), b.f64("f_elem"))
    assert abs(a.scalar("ala_n") - -12.116887) < 1e-5      # thoout.dat of the reference's hfbtho_main run


def test_front_end_basis_and_orthonormality(tmp_path):
    stage_point("S40_SKOP_6sh", "GT-K0", 0, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    assert (p.iscalar("dqp"), p.iscalar("nb"), p.iscalar("nghl"), p.iscalar("nxy")) == (168, 26, 400, 1640)
    wf = p.table("wf")
    ns, nl, nz = p.i32("ns"), p.i32("nl"), p.i32("nz")
    g = wf.T @ wf
    # quadrature-orthonormal within equal (Lambda, z-parity); the z mesh covers z>0 only (reflection symmetry)
    # and the table carries sqrt(weights)
    same = (nl[:, None] == nl[None, :]) & ((nz[:, None] % 2) == (nz[None, :] % 2)) & (ns[:, None] == ns[None, :])
    assert np.abs((g - np.eye(168))[same]).max() < 1e-12
    # spin-up states first inside every block
    db, nsu = p.i32("db"), p.i32("num_spin_up")
    o = 0
    for d, u in zip(db, nsu):
        assert (ns[o:o + u] == 1).all() and (ns[o + u:o + d] == -1).all()
        o += d
    # U^T U + V^T V = 1 on the active quasiparticles (columns above the pairing window are zeroed)
    for t in "np":
        U, V, E = p.f64("U" + t), p.f64("V" + t), p.f64("E" + t)
        o = q = 0
        for d in db:
            u, v = U[o:o + d * d].reshape(d, d, order="F"), V[o:o + d * d].reshape(d, d, order="F")
            nrm = (u * u).sum(0) + (v * v).sum(0)
            act = E[q:q + d] != 0
            assert np.abs(nrm[act] - 1).max() < 1e-12 and (nrm[~act] == 0).all()
            o += d * d
            q += d
    # particle number from the normalised coordinate-space densities
    w = 1.0 / p.f64("wdcori")
    assert abs((p.f64("rho_n") * w).sum() - 24) < 1e-9 and abs((p.f64("rho_p") * w).sum() - 16) < 1e-9


def test_finite_temperature_setup(tmp_path):
    """Finite-temperature HFB solution (T = 0.8 MeV fixture made with the reference's executables): thermal occupations
    re-made from the quasiparticle energies (hfbtho_solution.f90:364-388), zeroed outside the pairing window; the largest
    ones against the reference's own log header (pnfam_txtoutput.f90:179-186); densities weighted with them integrate
    to N and Z (DENSIT, hfbtho_solver.f90:4535-4545)."""
    stage_point("Gd162_finiteT_6sh", "GT-K0", 0, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    assert p.iscalar("ft_active") == 1 and p.iscalar("blo_active") == 0 and p.iscalar("statistical") == 1
    assert p.scalar("ft_temp") == 0.8
    fp, fn, Ep, En = p.f64("qp_fp"), p.f64("qp_fn"), p.f64("Ep"), p.f64("En")
    ref = open(os.path.join(GOLDEN, "Gd162_finiteT_6sh", "reference_stdout.txt")).read()
    import re
    mp = re.search(r"max\. f_p \.+:\s+(\S+) \(Ep=(\S+)MeV\)", ref)
    mn = re.search(r"max\. f_n \.+:\s+(\S+) \(En=(\S+)MeV\)", ref)
    assert abs(fp.max() - float(mp.group(1))) < 6e-5 and abs(Ep[fp.argmax()] - float(mp.group(2))) < 6e-5
    assert abs(fn.max() - float(mn.group(1))) < 6e-5 and abs(En[fn.argmax()] - float(mn.group(2))) < 6e-5
    for f, E in ((fp, Ep), (fn, En)):
        act = E != 0
        assert (f[~act] == 0).all() and np.allclose(f[act], 0.5 * (1 - np.tanh(0.5 * E[act] / 0.8)), rtol=0, atol=1e-15)
    w = 1.0 / p.f64("wdcori")
    assert abs((p.f64("rho_n") * w).sum() - 98) < 1e-9 and abs((p.f64("rho_p") * w).sum() - 64) < 1e-9


@pytest.mark.parametrize("case,op,idx", [("S40_SKOP_6sh", "GT-K0", 10), ("S40_GT_All", "RS2-K2", 7), ("Gd162_GT_open_6sh", "GT-K1", 40),
                                         ("Gd163_blocked_6sh", "GT-K1", 0), ("Gd162_finiteT_6sh", "RS1-K1", 0)])
def test_fused_transform_jobs_equal_the_block_task_list(case, op, idx, tmp_path):
    """The fused transform kernel works on jobs regrouped by associativity (products shared by two output blocks are
    formed once).  Host-only self-check of the C ABI library: both forms evaluated on the CPU with random operands
    give the same block matrices, and the regrouping needs fewer products than triprod_bbm's term list
    (pnfam_type_bbm.f90:428-552): 8 -> 7 / 6 per block row (forward / backward), 16 -> 10 / 8 with the P,Q quadrants."""
    from pynfam_b200 import gpu
    stage_point(case, op, idx, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    L = gpu.lib()
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    L.pnfam_b200_check_transform_plan.argtypes = [ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_int, dp, ctypes.c_char_p, ctypes.c_int]
    db, r2c = np.ascontiguousarray(p.i32("db")), np.ascontiguousarray(p.i32("f_ir2c"))
    out, err = (ctypes.c_double * 8)(), ctypes.create_string_buffer(512)
    stat = int(p.iscalar("statistical"))
    rc = L.pnfam_b200_check_transform_plan(len(db), db.ctypes.data_as(ip), r2c.ctypes.data_as(ip), stat, int(p.iscalar("beta_minus")), out, err, 512)
    assert rc == 0, err.value
    for d in (0, 1):
        ntasks, njobs, ratio, diff = out[4 * d:4 * d + 4]
        assert njobs <= ntasks and diff < 1e-13
        assert ratio <= (0.63 if stat else 0.88)


def test_couplings_match_reference_header(tmp_path):
    """The .dat header of the golden point prints the couplings to 9 decimals."""
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path))
    p = host.Problem(str(tmp_path), "x.in")
    want = dict(cr0=246.942585311, crr=-198.724045856, cs0=128.279, csr=0.0, cdrho=-32.162034342, ctau=-4.156239521,
                cj=4.156239521, crdj=41.4444, csdj=41.4444, ctj0=3.057291667, ctj1=4.5859375, ctj2=9.171875,
                ct=-9.171875, cf=0.0, cpair0=-33.027032645, cpairr=103.100736927, cspair0=-43.294, cspairr=135.151206361)
    for k, v in want.items():
        assert abs(p.scalar(k) - v) < 6e-10, (k, p.scalar(k), v)


def test_bad_inputs_fail_cleanly(tmp_path):
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path),
                patch=lambda s: s.replace("operator_name = 'GT'", "operator_name = 'XYZ'"))
    with pytest.raises(host.PnfamError, match="Unknown operator"):
        host.Problem(str(tmp_path), "x.in")
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path),
                patch=lambda s: s.replace("operator_k = 0", "operator_k = 2"))
    with pytest.raises(host.PnfamError, match="K out of range"):
        host.Problem(str(tmp_path), "x.in")
    with pytest.raises(host.PnfamError):
        host.Problem(str(tmp_path / "nowhere"), "x.in")
    # couplings from a file (pnfam_interaction.f90:303-331): a missing file and a gauge-invariance violation are errors
    stage_point("S40_custom_interaction", "GT-K0", 0, str(tmp_path / "c"))
    os.remove(str(tmp_path / "c" / "custom_edf.dat"))
    with pytest.raises(host.PnfamError, match="could not open interaction file"):
        host.Problem(str(tmp_path / "c"), "x.in")
    stage_point("S40_custom_interaction", "GT-K0", 0, str(tmp_path / "d"))
    f = tmp_path / "d" / "custom_edf.dat"
    f.write_text(re.sub(r"cj\s*=\s*\S+", "cj = 1.0", f.read_text()))
    with pytest.raises(host.PnfamError, match="gauge invariance"):
        host.Problem(str(tmp_path / "d"), "x.in")


def test_drop_in_exe_without_gpu_reports_failure_the_reference_way(tmp_path):
    """Without a CUDA device the drop-in must not fall back to a CPU path: like the reference on an error it
    exits 0, keeps stderr silent and emits no result table, which is how pynfam detects a failed task
    (pynfam/fortran/pnfam_run.py:155-164)."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = os.path.join(ROOT, "pynfam_b200", "bin", "pnfam_main.x")
    assert os.path.isfile(exe)
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path), name="GT-K0.in")
    r = subprocess.run([exe, "GT-K0.in"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stderr.strip() == ""
    assert "no CUDA device" in r.stdout
    assert "Strength" not in refrun.parse_dat(r.stdout)["rows"]
    assert not os.path.isfile(str(tmp_path / "GT-K0.dat"))


def test_contour_exe_without_gpu_reports_failure_the_reference_way(tmp_path):
    """contour_main.x (drop-in for exes/pnfam/contour_prog.f90): namelists parsed, operator list built, and without a
    CUDA device no summary file appears (no CPU fallback); exit 0 and silent stderr like the reference's `stop`."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = os.path.join(ROOT, "pynfam_b200", "bin", "contour_main.x")
    assert os.path.isfile(exe)
    stage_point("S40_SKOP_6sh", "GT-K0", 10, str(tmp_path), name="pnfam_NAMELIST.dat")
    (tmp_path / "pnfam_CONTOUR.dat").write_text(
        "&ctr_general\n fam_mode = 'STR'\n fam_input_filename = 'pnfam_NAMELIST.dat'\n/\n"
        "&ctr_extfield\n operator_groups = '0+', '1+', '0-', '1-', '2-'\n operator_active = 0, 1, 0, 0, 0\n/\n"
        "&str_parameters\n energy_start = 0.0\n energy_step = 1.0\n nr_points = 3\n half_width = 0.25\n/\n")
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stderr.strip() == ""
    assert "no CUDA device" in r.stdout
    assert not os.path.isfile(str(tmp_path / "GT-K0.out"))
    r = subprocess.run([exe, "a", "b"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "More than one command-line argument" in r.stdout
    (tmp_path / "bad.dat").write_text("&ctr_general\n fam_mode = 'FINDMAX'\n/\n")
    r = subprocess.run([exe, "bad.dat"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "not available" in r.stdout


# ---- the Fortran binding source against the C header ------------------------------------------------------
_CTYPES = {"int32_t": ("integer(c_int32_t)", 4), "int64_t": ("integer(c_int64_t)", 8), "double": ("real(c_double)", 8),
           "ptr": ("type(c_ptr)", 8)}


def _c_structs():
    """{struct name: [(kind, field)]} of the typedef'd structs of include/pnfam_b200.h, in declaration order."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "pnfam_b200.h")).read(), flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", hdr, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(.*)", decl)
            base, rest = m.group(2), m.group(3)
            for item in rest.split(","):
                item = item.strip()
                nm = re.sub(r"[\s\*]|const", "", item)
                fields.append(("ptr" if "*" in item else base, nm))
        out[name] = fields
    return out


def _f90_types():
    src = open(os.path.join(ROOT, "include", "pnfam_b200_binding.f90")).read()
    out = {}
    for name, body in re.findall(r"type,\s*bind\(C\)\s*::\s*(\w+)\n(.*?)\n\s*end type", src, flags=re.S):
        fields = []
        for ln in body.split("\n"):
            ln = ln.split("!")[0].strip()
            if not ln:
                continue
            kind, names = [x.strip() for x in ln.split("::")]
            fields += [(kind, n.strip()) for n in names.split(",")]
        out[name] = fields
    return out


def test_fortran_binding_matches_the_c_structs(tmp_path):
    """include/pnfam_b200_binding.f90 declares every struct of the header field for field (name, kind, order), and the
    byte offsets a Fortran bind(C) type gets (natural alignment, the C interoperability rule) equal offsetof() of the
    C structs as compiled here -- a caller built from the binding passes exactly what the library reads."""
    import subprocess
    cs, fs = _c_structs(), _f90_types()
    assert set(cs) == set(fs) and len(cs) == 5
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "pnfam_b200.h"', 'int main(void){']
    for name, fields in cs.items():
        assert [n.lower() for _, n in fields] == [n.lower() for _, n in fs[name]], name
        for (ck, cn), (fk, fn) in zip(fields, fs[name]):
            assert _CTYPES[ck][0] == fk, (name, cn, ck, fk)
            prog.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, cn, name, cn))
        prog.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (name, name))
    prog.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(prog))
    exe = str(tmp_path / "probe")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)], check=True)
    got = dict(ln.split() for ln in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, fields in fs.items():
        off, amax = 0, 1
        for kind, fn in fields:
            size = [v[1] for v in _CTYPES.values() if v[0] == kind][0]
            off = (off + size - 1) // size * size
            cname = [n for _, n in cs[name] if n.lower() == fn.lower()][0]
            assert int(got["%s.%s" % (name, cname)]) == off, (name, fn)
            off += size
            amax = max(amax, size)
        assert int(got[name + ".sizeof"]) == (off + amax - 1) // amax * amax, name
    # every entry point of section 2 is bound
    f90 = open(os.path.join(ROOT, "include", "pnfam_b200_binding.f90")).read()
    for sym in _declared_symbols():
        if sym.startswith("pnfam_b200_"):
            assert 'name="%s"' % sym in f90, sym
