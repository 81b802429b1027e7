#!/usr/bin/env python
"""Parity numbers on the GPU box (development tool; the asserted versions live in tests/test_gpu_production.py):

  python scripts/parity_probe.py prod  <case dir name> <points file>     # batched solve of every point of a fixture
  python scripts/parity_probe.py loose                                   # ill-conditioned 6-shell points: GPU vs the
                                                                         # reference run here vs the reference's golden file
Prints one line per point and the maxima; writes gpurun_out/parity_<tag>.json."""
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN  # noqa: E402
from pynfam_b200 import gpu, host  # noqa: E402


def omega_of(nml):
    return complex(float(re.search(r"real_eqrpa\s*=\s*(\S+)", nml).group(1)), float(re.search(r"imag_eqrpa\s*=\s*(\S+)", nml).group(1)))


def stage(case, nml, wd, name):
    import shutil
    for f in os.listdir(os.path.join(GOLDEN, case)):
        if f.startswith("hfbtho_") or f.endswith(".tbc"):
            if not os.path.isfile(os.path.join(wd, f)):
                shutil.copy(os.path.join(GOLDEN, case, f), wd)
    open(os.path.join(wd, name), "w").write(re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", nml))


def prod(case, fname):
    pts = json.load(open(os.path.join(GOLDEN, case, fname)))["points"]
    wd = tempfile.mkdtemp()
    base, ctx, out = None, None, []
    for op, lst in pts.items():
        stage(case, lst[0]["namelist"], wd, op + ".in")
        p = host.Problem(wd, op + ".in", share_nucleus_with=base)
        if base is None:
            base, ctx = p, gpu.Context(p)
        r = ctx.solve(p, omegas=[omega_of(pt["namelist"]) for pt in lst])
        for i, pt in enumerate(lst):
            worst = 0.0
            for k, lab in enumerate(["Strength"] + r["labels"][1:]):
                if lab in pt["rows"]:
                    g = complex(float(pt["rows"][lab][0]), float(pt["rows"][lab][1]))
                    if abs(g) > 0:
                        worst = max(worst, abs(r["strength"][i, k] - g) / abs(g))
            s = r["strength"][i, 0]
            g = complex(float(pt["rows"]["Strength"][0]), float(pt["rows"]["Strength"][1]))
            rec = {"op": op, "i": i, "omega": [omega_of(pt["namelist"]).real, omega_of(pt["namelist"]).imag], "iters_ref": pt["iters"],
                   "iters_gpu": int(r["iters"][i]), "conv_gpu": int(r["conv"][i]), "rel_S": abs(s - g) / abs(g), "rel_rows": worst}
            out.append(rec)
            print("%-7s %2d w=%9.4f%+9.4fi it %3d/%3d  rel S %.2e  rows %.2e" % (op, i, rec["omega"][0], rec["omega"][1], rec["iters_gpu"],
                  rec["iters_ref"], rec["rel_S"], worst), flush=True)
        print(op, "solve stats", {k: r["stats"][k] for k in ("seconds_device", "iterations", "batch_slots", "lock_steps")}, flush=True)
    print("MAX rel S %.2e rows %.2e" % (max(o["rel_S"] for o in out), max(o["rel_rows"] for o in out)))
    return out


def loose():
    d = json.load(open(os.path.join(GOLDEN, "loose_points_6sh.json")))["cases"]
    out = []
    for case, lst in d.items():
        allpts = json.load(open(os.path.join(GOLDEN, case, "points.json")))["points"]
        wd = tempfile.mkdtemp()
        base, ctx = None, None
        byop = {}
        for r_ in lst:
            byop.setdefault(r_["name"], []).append(r_)
        for op, rs in byop.items():
            nml = {p["point"]: p["namelist"] for p in allpts[op]}
            stage(case, nml[rs[0]["point"]], wd, op + ".in")
            p = host.Problem(wd, op + ".in", share_nucleus_with=base)
            if base is None:
                base, ctx = p, gpu.Context(p)
            r = ctx.solve(p, omegas=[omega_of(nml[x["point"]]) for x in rs])
            for i, x in enumerate(rs):
                here = complex(float(x["here"][0]), float(x["here"][1]))
                gold = complex(float(x["golden"][0]), float(x["golden"][1]))
                s = r["strength"][i, 0]
                rec = {"case": case, "op": op, "point": x["point"], "iters_gpu": int(r["iters"][i]), "iters_here": x["here_iters"],
                       "iters_golden": x["golden_iters"], "gpu_vs_here": abs(s - here) / abs(here), "gpu_vs_golden": abs(s - gold) / abs(gold),
                       "here_vs_golden": x["rel"], "im": omega_of(nml[x["point"]]).imag}
                out.append(rec)
        sel = [o for o in out if o["case"] == case]
        print("%-22s %3d points: max gpu-vs-here %.2e  gpu-vs-golden %.2e  here-vs-golden %.2e ; iteration counts equal to here: %d, to golden: %d"
              % (case, len(sel), max(o["gpu_vs_here"] for o in sel), max(o["gpu_vs_golden"] for o in sel), max(o["here_vs_golden"] for o in sel),
                 sum(o["iters_gpu"] == o["iters_here"] for o in sel), sum(o["iters_gpu"] == o["iters_golden"] for o in sel)), flush=True)
    for o in sorted(out, key=lambda o: -o["gpu_vs_here"])[:15]:
        print(o)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if sys.argv[1] == "prod":
        res = prod(sys.argv[2], sys.argv[3])
        tag = sys.argv[2] + "_" + sys.argv[3].replace(".json", "")
    else:
        res = loose()
        tag = "loose_6sh"
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_%s.json" % tag), "w"), indent=0)
