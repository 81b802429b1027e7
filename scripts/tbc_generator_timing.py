#!/usr/bin/env python
"""Wall time of the host set-up with the full-FAM two-body-current field generator (csrc/host/tbc_generator.cpp) for
162Gd at the basis sizes of BASELINE.json configs[4]: run directory without a .tbc file, mode 111100.
The reference quotes 1.5-2.5 h per K at 16 shells on 44 OpenMP threads (exes/pnfam/README_2bc.md:113-118).
usage: tbc_generator_timing.py [shells ...]  ->  profiles/r02_tbc_generator.json"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_gd162_16sh import FAM  # noqa: E402

CHILD = """
import sys, time
sys.path.insert(0, %r)
from pynfam_b200 import host
t = time.time(); p = host.Problem(sys.argv[1], sys.argv[2]); print("SETUP_S", time.time() - t, p.iscalar("nxy"))
"""


def main():
    shells = [int(a) for a in sys.argv[1:]] or [12, 16, 20]
    out_path = os.path.join(ROOT, "profiles", "r02_tbc_generator.json")
    res = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    res["what"] = ("host set-up of a GT field with full-FAM two-body currents, no .tbc file in the run directory "
                   "(HFB reconstruction + field generator + cache write), 162Gd SkO'")
    res["threads"] = os.cpu_count()
    res["reference"] = "1.5-2.5 h per K at 16 shells on 44 OpenMP threads (exes/pnfam/README_2bc.md:113-118)"
    for sh in shells:
        for k, usep in ((0, False), (1, True)):
            wd = tempfile.mkdtemp()
            g = os.path.join(ROOT, "tests", "golden", "Gd162_SKOP_%dsh" % sh)
            for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
                shutil.copy(os.path.join(g, f), wd)
            name = "GT-K%d" % k
            nml = FAM.format(name=name, re="2.0", im="1.0", op="GT", k=k, max_iter=300)
            nml = nml.replace("two_body_current_mode = 0", "two_body_current_mode = 111100")
            if usep:
                nml = nml.replace("two_body_current_usep = .false.", "two_body_current_usep = .true.")
            open(os.path.join(wd, name + ".in"), "w").write(nml)
            env = dict(os.environ, PNFAM_B200_NO_CACHE="1", PNFAM_B200_SETUP_TIMING="1")
            t = time.time()
            r = subprocess.run([sys.executable, "-c", CHILD % ROOT, wd, name + ".in"], env=env, capture_output=True, text=True)
            wall = time.time() - t
            assert r.returncode == 0, r.stderr[-2000:]
            phases = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"\[setup\]\s+2BC (.*?)\s+([0-9.]+) s", r.stderr)}
            m = re.search(r"SETUP_S ([0-9.e+-]+) (\d+)", r.stdout)
            key = "%dsh_K%d%s" % (sh, k, "_usep" if usep else "")
            res[key] = {"setup_s": float(m.group(1)), "nxy": int(m.group(2)), "generator_phases_s": phases, "process_wall_s": wall,
                        "tbc_bytes": os.path.getsize(os.path.join(wd, name + ".tbc"))}
            print(key, res[key], flush=True)
            json.dump(res, open(out_path, "w"), indent=1)
            shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
