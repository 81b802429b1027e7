#!/usr/bin/env python
"""Copy the reference's own strength-assembly outputs (pynfam's OP.out text summary and OP.out.ctr binary, written by
famStrength.writeStrengthOut / writeCtrBinary, pynfam/strength/fam_strength.py:456-568) for a few operators of
tests/pynfam_test_S40 and tests/S40_GT_All into tests/golden/<case>/fam_soln/.  Data files of the reference's test tree, not source.
Run ONCE in the build container (where /root/reference exists); the results are committed."""
import os
import shutil

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
# (tree, golden case, operators).  The forbidden operators come from the 2024 tree: the RS1xP cross-term of the older
# pynfam_test_S40 tree was produced by an older source (the live reference binary agrees with the newer tree).
SETS = [("pynfam_test_S40/000000/fam_soln", "S40_SKOP_6sh", ("GT-K0", "GT-K1")),
        ("S40_GT_All/000000/fam_soln", "S40_GT_All", ("RS1-K1", "PS0-K0"))]
for tree, case, ops in SETS:
    dst = os.path.join(HERE, case, "fam_soln")
    os.makedirs(dst, exist_ok=True)
    for op in ops:
        for ext in (".out", ".out.ctr"):
            shutil.copy(os.path.join(REF, tree, op + ext), os.path.join(dst, op + ext))
            os.chmod(os.path.join(dst, op + ext), 0o644)
    print(case, sorted(os.listdir(dst)))
