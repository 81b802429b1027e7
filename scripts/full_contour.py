#!/usr/bin/env python
"""Full beta-decay contour of one nucleus on the GPUs of one box (BASELINE.json configs[3]): all 14 (operator, K)
of the allowed + first-forbidden set x the computed half of pynfam's 60-node CIRCLE contour, cross-terms on, sharded
over the ranks by pynfam_b200.strength.run_contours_sharded (one process per GPU, NCCL only for the final all_reduce
of the strengths).  Writes OP.out / OP.out.ctr for every operator into --dest and prints one JSON line.

  python scripts/full_contour.py --shells 20                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \\
      scripts/full_contour.py --shells 20                         # 8 GPUs
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    # kept for the command lines quoted in DESIGN.md / profiles: the implementation lives in bench.py --full-contour
    if "--full-contour" not in sys.argv:
        sys.argv.append("--full-contour")
    if "--shells" not in sys.argv:
        sys.argv += ["--shells", "20"]
    bench.main()
