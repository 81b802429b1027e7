#!/usr/bin/env python
"""Development probe (GPU box): fused vs two-phase transform kernels -- the same solves at fixed iteration counts in two
subprocesses (PNFAM_B200_TRANSFORM_2PHASE toggled) and the relative differences of the strengths."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CASES = [("S40_SKOP_6sh", "GT-K0", 10), ("Gd162_GT_open_6sh", "GT-K1", 40), ("S40_GT_All", "RS2-K2", 7), ("Gd163_blocked_6sh", "GT-K1", 0),
         ("Gd162_finiteT_6sh", "RS1-K1", 0), ("S40_All_GT2bc", "GT-K0", 27)]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from conftest import stage_point
    from pynfam_b200 import gpu, host
    out = {}
    for case, op, idx in CASES:
        wd = tempfile.mkdtemp()
        stage_point(case, op, idx, wd)
        p = host.Problem(wd, "x.in")
        ctx = gpu.Context(p)
        for mi in (1, 3, 8, 300):
            r = ctx.solve(p, max_iter=mi)
            out["%s/%s/%d/%d" % (case, op, idx, mi)] = [[float(z.real), float(z.imag)] for z in r["strength"][0]] + [[int(r["iters"][0]), 0.0]]
    print(json.dumps(out))
    sys.exit(0)
res = []
for env in ({}, {"PNFAM_B200_TRANSFORM_2PHASE": "1"}):
    e = dict(os.environ); e.update(env)
    o = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
    if o.returncode != 0:
        print(o.stderr[-3000:]); sys.exit(1)
    res.append(json.loads(o.stdout.strip().splitlines()[-1]))
for k in res[0]:
    a, b = res[0][k], res[1][k]
    worst = max(abs(complex(*x) - complex(*y)) / max(abs(complex(*y)), 1e-300) for x, y in zip(a[:-1], b[:-1]))
    print("%-40s iters %3d / %3d   fused vs two-phase: max rel %.2e" % (k, a[-1][0], b[-1][0], worst))
