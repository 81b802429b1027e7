#!/usr/bin/env python
"""Level-0 measurement (GPU box): the path pynfam drives unchanged -- one `pnfam_main.x <namelist>` PROCESS per omega point
(pynfam/fortran/fortran_utils.py:222-247), each paying process start, CUDA context, the HFB reconstruction (or its cache) and
the solve -- next to the `&b200_batch` extension (all points of an operator in one launch).  162Gd, 16 shells, GT- K=0,
the first 8 nodes of bench.py's contour sweep.  Prints one JSON line."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

exe = os.path.join(ROOT, "pynfam_b200", "bin", "pnfam_main.x")
om = bench.sweep_contour(64)[::8][:8]
wd = tempfile.mkdtemp()
bench.stage(wd, om[0], 300)
nml = open(os.path.join(wd, "GT-K0.in")).read()


def run(text):
    open(os.path.join(wd, "GT-K0.in"), "w").write(text)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "GT-K0.in"], cwd=wd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0 and r.stderr.strip() == "" and os.path.isfile(os.path.join(wd, "GT-K0.dat")), r.stdout[-2000:]
    return dt


import re
per_point = []
for k, w in enumerate(om):                      # the first launch also writes the set-up cache, as a pynfam run would
    t = re.sub(r"real_eqrpa = \S+", "real_eqrpa = %r" % float(w.real), nml)
    t = re.sub(r"imag_eqrpa = \S+", "imag_eqrpa = %r" % float(w.imag), t)
    per_point.append(run(t))
batch = nml + "\n&b200_batch\n    real_eqrpa = %s\n    imag_eqrpa = %s\n/\n" % (", ".join(repr(float(w.real)) for w in om), ", ".join(repr(float(w.imag)) for w in om))
t_batch = run(batch)
print(json.dumps({"workload": "Gd162 SkO' 16 shells, GT- K=0, 8 omega points of the bench sweep, one pnfam_main.x process per point (pynfam's "
                              "unchanged launch path) vs one process with &b200_batch",
                  "seconds_per_process": per_point, "first_process_s (writes the set-up cache)": per_point[0],
                  "mean_later_process_s": sum(per_point[1:]) / (len(per_point) - 1), "omega_points_per_s_level0": (len(om) - 1) / sum(per_point[1:]),
                  "b200_batch_process_s": t_batch, "omega_points_per_s_batch": len(om) / t_batch}))
