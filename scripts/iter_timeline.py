"""Development probe (GPU box): wall time of every lock-step iteration of the bench solve and the number of active
omega points in it."""
import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from pynfam_b200 import host, gpu

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wd = tempfile.mkdtemp()
oms = bench.circle_contour(npts)
bench.stage(wd, oms[0], 300)
p = host.Problem(wd, "GT-K0.in")
ctx = gpu.Context(p)
ctx.solve(p, omegas=oms)
r = ctx.solve(p, omegas=oms, want_trace=True)
iters = np.array(r["iters"])
tr = r["trace"]            # [P][max_iter+1][4], column 3 = dt of the lock-step iteration
nit = iters.max()
tot = 0.0
print("iteration  active  ms   ms/point")
for it in range(1, nit + 1):
    act = int((iters >= it).sum())
    dt = max(tr[p_, it, 3] for p_ in range(npts) if iters[p_] >= it) * 1e3
    tot += dt
    print(f"{it:4d} {act:4d} {dt:8.2f} {dt/act:8.2f}")
print("total ms", tot, "point-iterations", iters.sum(), "stats device s", r["stats"]["seconds_device"])
