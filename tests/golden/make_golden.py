#!/usr/bin/env python
"""Generate tests/golden/<case>/ from the reference's own golden output trees.

Run ONCE in the build container (where /root/reference exists); the results are committed.
Each case directory gets
  hfbtho_NAMELIST.dat, hfbtho_output.hel   -- verbatim copies of the reference fixture's INPUT data
                                             files (hfb_soln/), not source code
  points.json                              -- for every per-point run stored in fam_meta/OP.tar:
                                             the namelist text (OP.in) and the known answers parsed
                                             from OP.dat (19-digit result table, iteration trace).
The source of every number is tests/<tree>/000000/fam_soln/fam_meta/*.tar in mld1812/pynfam.
"""
import io
import json
import os
import shutil
import sys
import tarfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.refrun import parse_dat  # noqa: E402

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # case name: (tree relative to REF, operators to keep (None = all present))
    "S40_SKOP_6sh": ("pynfam_test_S40/000000", None),
    "S40_GT_All": ("S40_GT_All/000000", None),
    "Gd162_GT_open_6sh": ("Gd162_GT_open/000000", None),
    "Gd162_1-_closed_6sh": ("Gd162 closed tests/1-/000000", None),
    "Gd162_0-_closed_6sh": ("Gd162 closed tests/0-/000000", None),
    # full-FAM two-body currents via hfb_soln/*.tbc.  Only the GT operators: the forbidden-operator results of
    # this 2023 tree were made with an older source whose mode digit 4 = 1 also corrected RS*; the shipped
    # binary/source (pnfam_extfield.f90:430-438: use_2bc(4) >= 2) does not, and the live binary agrees with us.
    "S40_All_GT2bc": ("S40_All_GT2bc/000000", ["GT-K0", "GT-K1"]),
}


def main():
    for case, (tree, ops) in CASES.items():
        src = os.path.join(REF, tree)
        if not os.path.isdir(src):
            print("skip (absent):", src)
            continue
        dst = os.path.join(HERE, case)
        os.makedirs(dst, exist_ok=True)
        for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
            shutil.copy(os.path.join(src, "hfb_soln", f), dst)
        for f in os.listdir(os.path.join(src, "hfb_soln")):
            if f.endswith(".tbc"):   # cached two-body-current external field (input data of the run)
                shutil.copy(os.path.join(src, "hfb_soln", f), dst)
        points = {}
        meta = os.path.join(src, "fam_soln", "fam_meta")
        for tarname in sorted(os.listdir(meta)):
            if not tarname.endswith(".tar"):
                continue
            op = tarname[:-4]
            if ops and op not in ops:
                continue
            try:
                tf = tarfile.open(os.path.join(meta, tarname))
                members = {m.name: m for m in tf.getmembers()}
            except Exception as e:  # missing large blobs are stubs
                print("unreadable", tarname, e)
                continue
            pts = []
            for name in sorted(members):
                if not name.endswith(".in"):
                    continue
                datname = name[:-3] + ".dat"
                if datname not in members:
                    continue
                nml = tf.extractfile(members[name]).read().decode()
                dat = parse_dat(tf.extractfile(members[datname]).read().decode())
                if "Strength" not in dat["rows"]:
                    continue
                pts.append({
                    "point": name.split("/")[1],
                    "namelist": nml,
                    "rows": {k: [repr(v.real), repr(v.imag)] for k, v in dat["rows"].items()},
                    "iters": dat["iters"], "conv": dat["conv"],
                    "trace": [[t[0], t[1], t[2], t[3], t[4]] for t in dat["trace"]],
                    "header": dat["header"],
                })
            if pts:
                points[op] = pts
        json.dump({"source": "mld1812/pynfam tests/" + tree, "points": points},
                  open(os.path.join(dst, "points.json"), "w"), indent=0)
        print(case, {k: len(v) for k, v in points.items()},
              os.path.getsize(os.path.join(dst, "points.json")) // 1024, "KiB")


if __name__ == "__main__":
    main()
