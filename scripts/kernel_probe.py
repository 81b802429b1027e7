"""Development timing probe (GPU box): per-launch time of the density / projection kernels on the bench workload
for a fixed number of iterations.  Usage: python scripts/kernel_probe.py [points] [iters]"""
import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from pynfam_b200 import host, gpu

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wd = tempfile.mkdtemp()
oms = bench.circle_contour(npts)
bench.stage(wd, oms[0], iters)
p = host.Problem(wd, "GT-K0.in")
ctx = gpu.Context(p)
for rep in range(2):
    r = ctx.solve(p, omegas=oms, max_iter=iters)
st = r["stats"]
print("PROJ_DEBUG=%s DENS_DEBUG=%s points %d: density %.3f ms/launch, projection %.3f ms/launch, device total %.1f ms" % (
    os.environ.get("PNFAM_B200_PROJ_DEBUG", "0"), os.environ.get("PNFAM_B200_DENS_DEBUG", "0"), npts, 1e3 * st["seconds_density"] / max(1, st["launches_density"]),
    1e3 * st["seconds_projection"] / max(1, st["launches_projection"]), 1e3 * st["seconds_device"]))
