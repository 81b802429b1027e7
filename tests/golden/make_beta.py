#!/usr/bin/env python
"""Golden data for the integrated beta-decay rate (tests/pynfam_test_S40/000000/beta_soln): the allowed-channel rows of
the reference's beta.out and the matching columns of its phase-space weighted shape factor at every contour point
(shapefactor_im.out, beta_meta/shapefactor_re.out; pynfam/strength/shape_factor.py:293-357, 956-1031).  For an allowed
channel the weighted shape factor is (phase-space function of the complex energy) x (one FAM strength), so
shape factor / strength gives the reference's own integration weights at the contour points without restating
phase_space.py -- tests/test_strength.py and the GPU parity test use them to check the integrated rate.
Run ONCE in the build container (where /root/reference exists); the JSON is committed."""
import json
import os

SRC = "/root/reference/tests/pynfam_test_S40/000000/beta_soln"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "S40_SKOP_6sh", "beta_soln.json")
COLS = ["Allowed-Fermi", "Allowed-GT_K=0", "Allowed-GT_K=1"]


def table(path):
    lines = [ln for ln in open(path).read().split("\n") if ln.strip() and not ln.startswith("#")]
    head = lines[0].split()
    rows = [ln.split()[1:] for ln in lines[1:]]
    return {h: [r[i] for r in rows] for i, h in enumerate(head)}


im = table(os.path.join(SRC, "shapefactor_im.out"))
re = table(os.path.join(SRC, "beta_meta", "shapefactor_re.out"))
rates = {}
for ln in open(os.path.join(SRC, "beta.out")):
    t = ln.split()
    if t and t[0] in COLS:
        rates[t[0]] = {"rate": t[1], "halflife": t[2]}
out = {"source": "mld1812/pynfam tests/pynfam_test_S40/000000/beta_soln (beta.out, shapefactor_im.out, beta_meta/shapefactor_re.out)",
       "rates": rates,
       "shape_factor": {c: {"re": re[c], "im": im[c]} for c in COLS}}
json.dump(out, open(DST, "w"), indent=0)
print(DST, os.path.getsize(DST), rates)
