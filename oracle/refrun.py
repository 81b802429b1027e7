"""Test infrastructure -- NOT part of the product (only tests/, smoke() and bench.py's reference/cpu
baseline legs may import this).

Runs the UNMODIFIED prebuilt reference executables that ``make -C oracle`` copies into
``oracle/_ref/`` (``pnfam_main.x`` = pnFAM 2.00, ``hfbtho_main`` = HFBTHO v4), through the runtime
shim described in ``oracle/Makefile``.  Also holds the small namelist writer/reader and the ``.dat``
result parser that the reference drives through ``f90nml``/pandas
(pynfam/fortran/pnfam_run.py:208-270, pynfam/outputs/pnfam_parser.py:44-132), restated without
those packages (neither is installed in this image).
"""
import os
import re
import shutil
import subprocess
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _pylibs():
    return os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")


OPENBLAS = "libopenblasp-r0-59ffcd50.3.15.so"


def available():
    return (os.path.isfile(os.path.join(REF, "pnfam_main.x"))
            and os.path.isfile(os.path.join(_pylibs(), OPENBLAS)))


def ensure_built():
    """(Re)create the shim libraries if missing (symlinks do not always survive a snapshot)."""
    lib = os.path.join(REF, "lib")
    need = not os.path.isfile(os.path.join(lib, "libgfortran.so.5")) or \
        not os.path.exists(os.path.join(lib, "libblas.so.3"))
    if need:
        subprocess.run(["make", "-C", HERE, "-s", "REF=/nonexistent",
                        "_ref/lib/libgfortran.so.5", "_ref/lib/libblas.so.3"], check=True,
                       stdout=subprocess.DEVNULL)
    return available()


def _env(threads, tap=None):
    env = dict(os.environ)
    pylibs = _pylibs()
    libdir = os.path.join(REF, "tap" if tap else "lib")
    env["LD_LIBRARY_PATH"] = ":".join([libdir, os.path.join(REF, "lib"), pylibs])
    env["OMP_NUM_THREADS"] = str(threads)
    env["OPENBLAS_NUM_THREADS"] = str(threads)
    if tap:
        env["PNFAM_TAP"] = tap
        env["PNFAM_TAP_REAL"] = os.path.join(pylibs, OPENBLAS)
    return env


# ------------------------------------------------------------------------------------------------
# namelists
# ------------------------------------------------------------------------------------------------
def fmt_value(v):
    if v is None:
        return ","
    if isinstance(v, bool):
        return ".true." if v else ".false."
    if isinstance(v, str):
        return "'%s'" % v
    if isinstance(v, (list, tuple)):
        return ", ".join(fmt_value(x) for x in v)
    if isinstance(v, float):
        return repr(v)
    return str(v)


def write_namelist(path, groups):
    """groups: ordered dict {group: {key: value}} -> Fortran namelist file (f90nml style)."""
    with open(path, "w") as f:
        for g, kv in groups.items():
            f.write("&%s\n" % g)
            for k, v in kv.items():
                f.write("    %s = %s\n" % (k, fmt_value(v)))
            f.write("/\n\n")


def _parse_scalar(tok):
    t = tok.strip()
    if t == "":
        return None
    tl = t.lower()
    if tl in (".true.", "t", ".t."):
        return True
    if tl in (".false.", "f", ".f."):
        return False
    if t[0] in "'\"":
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(tl.replace("d", "e"))
    except ValueError:
        return t


def read_namelist(path):
    """Minimal Fortran-namelist reader: {group: {key: value-or-list}} (keys lower-cased)."""
    groups = {}
    cur = None
    for raw in open(path):
        line = raw.split("!")[0].strip()
        if not line:
            continue
        if line.startswith("&"):
            cur = line[1:].strip().lower()
            groups[cur] = {}
            continue
        if line.startswith("/"):
            cur = None
            continue
        if cur is None or "=" not in line:
            continue
        k, v = line.split("=", 1)
        v = v.strip()
        if v.endswith(",") and v != ",":
            v = v[:-1]
        if v == ",":
            groups[cur][k.strip().lower()] = None
            continue
        parts = [p for p in re.split(r",(?=(?:[^']*'[^']*')*[^']*$)", v)]
        vals = [_parse_scalar(p) for p in parts]
        groups[cur][k.strip().lower()] = vals[0] if len(vals) == 1 else vals
    return groups


# ------------------------------------------------------------------------------------------------
# .dat parser (contract of pynfam/outputs/pnfam_parser.py:44-132)
# ------------------------------------------------------------------------------------------------
def parse_dat(text):
    out = {"conv": None, "iters": None, "time_min": None, "version": None, "trace": [],
           "rows": {}, "header": {}}
    for line in text.splitlines():
        if line.startswith("#"):
            m = re.match(r"#\s+(\d+)([NLB])\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)", line)
            if m:
                out["trace"].append((int(m.group(1)), m.group(2), float(m.group(3)),
                                     float(m.group(4)), float(m.group(5))))
            if "iteration converged" in line:
                out["conv"] = True
                out["iters"] = int(re.search(r"after\s+(\d+)\s+steps", line).group(1))
            if "iteration interrupted" in line:
                out["conv"] = False
                out["iters"] = int(re.search(r"after\s+(\d+)\s+steps", line).group(1))
            if "Total CPU time" in line:
                out["time_min"] = float(line.split()[-2])
            if "Version:" in line:
                out["version"] = line.split()[-1]
            m = re.match(r"#\s+(Basis size|Number matrix blocks|Non-trivial HFB matrix elements|"
                         r"Non-trivial FAM matrix elements|Number shells):\s+(\d+)", line)
            if m:
                out["header"][m.group(1)] = int(m.group(2))
            continue
        toks = line.split()
        if len(toks) == 3 and toks[0] not in ("Real",):
            try:
                out["rows"][toks[0]] = complex(float(toks[1]), float(toks[2]))
            except ValueError:
                pass
    return out


# ------------------------------------------------------------------------------------------------
# running
# ------------------------------------------------------------------------------------------------
def run_pnfam(rundir, namelist, threads=1, tap=None, timeout=3600):
    """Run oracle/_ref/pnfam_main.x <namelist> in rundir.  Returns (parsed .dat, wall seconds,
    stdout).  rundir must hold hfbtho_NAMELIST.dat and hfbtho_output.hel
    (exes/pnfam/hfbtho_interface.f90:33)."""
    if not ensure_built():
        raise RuntimeError("oracle/_ref is not available (run `make -C oracle` where /root/reference exists)")
    t0 = time.time()
    p = subprocess.run([os.path.join(REF, "pnfam_main.x"), namelist], cwd=rundir, env=_env(threads, tap),
                       stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    return parse_dat(p.stdout), wall, p.stdout + p.stderr


def run_hfbtho(rundir, threads=1, timeout=3600):
    if not ensure_built():
        raise RuntimeError("oracle/_ref is not available")
    t0 = time.time()
    p = subprocess.run([os.path.join(REF, "hfbtho_main")], cwd=rundir, env=_env(threads),
                       stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=timeout)
    return p.stdout + p.stderr, time.time() - t0


def stage(rundir, hfb_dir, extra=()):
    os.makedirs(rundir, exist_ok=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel") + tuple(extra):
        shutil.copy(os.path.join(hfb_dir, f), rundir)
