#!/usr/bin/env python
"""Iteration counts of the N-GPU bench sweep (development tool): solves the 128*N points of bench.py's contour sweep on one
GPU and writes gpurun_out/sweep_iters_<N>.json = [[Re w, Im w, iterations], ...] for offline studies of the partition."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pynfam_b200 import gpu, host  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
om = bench.sweep_contour(128 * n)
wd = tempfile.mkdtemp()
bench.stage(wd, om[0], 300)
p = host.Problem(wd, "GT-K0.in")
ctx = gpu.Context(p)
r = ctx.solve(p, omegas=om)
json.dump([[float(w.real), float(w.imag), int(i)] for w, i in zip(om, r["iters"])], open(os.path.join(ROOT, "gpurun_out", "sweep_iters_%d.json" % n), "w"))
print("points", len(om), "iterations", int(r["iters"].sum()), "conv", int((r["conv"] > 0).sum()), "seconds", r["stats"]["seconds_device"])
