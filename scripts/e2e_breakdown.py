"""Where the end-to-end time of one bench step goes (dev tool): context creation, solve wall clock vs device time."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pynfam_b200 import gpu, host
om = bench.circle_contour(32)
wd = tempfile.mkdtemp(); bench.stage(wd, om[0], 300)
prob = host.Problem(wd, "GT-K0.in")
ctx = gpu.Context(prob)
for _ in range(2):
    ctx.solve(prob, omegas=om)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c2 = gpu.Context(prob); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = c2.solve(prob, omegas=om); torch.cuda.synchronize(); t2 = time.perf_counter()
    st = r["stats"]
    del c2; torch.cuda.synchronize(); t3 = time.perf_counter()
    print("ctx %.1f ms  solve wall %.1f ms (C ABI total %.1f, device loop %.1f)  destroy %.1f ms  iters %d" %
          (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * st["seconds_total"], 1e3 * st["seconds_device"], 1e3 * (t3 - t2), st["iterations"]))
