#!/usr/bin/env python
"""Full beta-decay contour of one nucleus on the GPUs of one box (BASELINE.json configs[3]): all 14 (operator, K)
of the allowed + first-forbidden set x the computed half of pynfam's 60-node CIRCLE contour, cross-terms on, sharded
over the ranks by pynfam_b200.strength.run_contours_sharded (one process per GPU, NCCL only for the final all_reduce
of the strengths).  Writes OP.out / OP.out.ctr for every operator into --dest and prints one JSON line.

  python scripts/full_contour.py --shells 20                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \\
      scripts/full_contour.py --shells 20                         # 8 GPUs
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (fixture staging + namelist of the bench workload)

# operators that share cross-term fields (same J^pi group and K) are neighbours, so a rank's contiguous run of tasks
# re-uses the fields cached per nucleus
OPERATORS = [("F-", 0), ("GT-", 0), ("GT-", 1), ("RS0-", 0), ("PS0-", 0), ("R-", 0), ("P-", 0), ("RS1-", 0), ("R-", 1), ("P-", 1),
             ("RS1-", 1), ("RS2-", 0), ("RS2-", 1), ("RS2-", 2)]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shells", type=int, default=20, choices=[16, 20])
    ap.add_argument("--nr-points", type=int, default=60)
    ap.add_argument("--emax", type=float, default=10.0)
    ap.add_argument("--dest", default=None)
    args = ap.parse_args()
    rank0, world0 = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun pins OMP_NUM_THREADS=1; the host set-up (HFB reconstruction, external fields) is OpenMP code
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world0))
    import numpy as np
    import torch
    import torch.distributed as dist
    from pynfam_b200.strength import famContour, run_contours_sharded
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)      # NCCL prints its banner on stdout from C code: fd 1 -> stderr until the JSON line
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bench.SHELLS = args.shells
    wd = tempfile.mkdtemp()
    bench.stage(wd, 1.0 + 1.0j, 300)
    dest = args.dest or tempfile.mkdtemp()
    contour = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": args.emax, "nr_points": args.nr_points})
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    fss = run_contours_sharded(wd, "GT-K0.in", OPERATORS, contour, dest=dest, dist=dist if world > 1 else None, device=local)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tm = run_contours_sharded.last_timing
    t = torch.tensor([wall, tm["host_setup_s"], tm["solve_s"], tm["gather_and_write_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        iters = int(sum(int(np.sum(f.iters)) for f in fss))
        npts = len(OPERATORS) * contour.nr_compute
        print(json.dumps({"workload": "Gd162 SkO' %d shells, %d (operator, K) x %d computed CIRCLE points, cross-terms on"
                                      % (args.shells, len(OPERATORS), contour.nr_compute),
                          "n_gpus": world, "seconds": float(t[0]), "omega_points": npts, "iterations": iters,
                          "omega_points_per_s": npts / float(t[0]), "iterations_per_s": iters / float(t[0]),
                          "max_over_ranks": {"host_setup_s": float(t[1]), "solve_s": float(t[2]), "gather_and_write_s": float(t[3])},
                          "host_threads_per_rank": int(os.environ["OMP_NUM_THREADS"]),
                          "all_converged": all(f.meta["Conv"] == "Yes" for f in fss),
                          "includes": "host set-up of the nucleus and of every operator on every rank, context creation, "
                                      "solves, all_reduce of the strengths, OP.out / OP.out.ctr written by rank 0",
                          "files": sorted(os.listdir(dest))[:4] + ["..."]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
