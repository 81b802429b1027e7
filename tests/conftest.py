import json
import os
import re
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (no-op when already built)."""
    lib = os.path.join(ROOT, "pynfam_b200", "lib")
    if not (os.path.isfile(os.path.join(lib, "libpnfam_host.so")) and os.path.isfile(os.path.join(lib, "libpnfam_b200.so"))):
        import __graft_entry__
        __graft_entry__.build()


def load_points(case):
    return json.load(open(os.path.join(GOLDEN, case, "points.json")))["points"]


def stage_point(case, op, idx, wd, name="x.in", patch=None):
    """Copy the case's HFB files into wd and write the point's namelist (old 3-digit 2BC mode 114 -> 0, which is
    what it meant for these results; the current reference source rejects 114)."""
    os.makedirs(wd, exist_ok=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        if not os.path.isfile(os.path.join(wd, f)):
            shutil.copy(os.path.join(GOLDEN, case, f), wd)
    for f in os.listdir(os.path.join(GOLDEN, case)):
        if (f.endswith(".tbc") or f == "custom_edf.dat") and not os.path.isfile(os.path.join(wd, f)):
            shutil.copy(os.path.join(GOLDEN, case, f), wd)
    pt = load_points(case)[op][idx]
    nml = re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", pt["namelist"])
    if patch:
        nml = patch(nml)
    with open(os.path.join(wd, name), "w") as f:
        f.write(nml)
    return pt


def gold_rows(pt):
    return {k: complex(float(v[0]), float(v[1])) for k, v in pt["rows"].items()}


def terminal_scatter(pt, back=3):
    """Largest relative change of S between consecutive iterations over the reference's last `back` steps (its own trace
    prints 10 digits).  Ill-conditioned points (>= 25 Broyden steps) stall near the stopping threshold: round-off differences
    between two correct implementations are amplified by ~1e7, the rule max|dX| < eps may then trigger a step or two
    apart, and S is only defined up to this movement."""
    tr = {t[0]: complex(t[3], t[4]) for t in pt["trace"]}
    last = max(tr)
    s = abs(tr[last])
    return max(abs(tr[k] - tr[k - 1]) / s for k in range(max(2, last - back + 1), last + 1) if k - 1 in tr)
