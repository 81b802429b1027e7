#!/bin/bash
# 2-GPU cycle: the two-devices-in-one-process test, Level-0 numbers, the bench at N=2, the 20-shell full contour at N=2.
tag=${1:-n2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu -x -k "two_contexts or sharded" 2>&1 | tail -5) > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
timeout 600 python scripts/level0.py > gpurun_out/level0_$tag.json 2> gpurun_out/level0_$tag.err; cat gpurun_out/level0_$tag.json; tail -3 gpurun_out/level0_$tag.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
  > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; tail -c 600 gpurun_out/bench_n2_$tag.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 \
  --full-contour --shells 20 > gpurun_out/full_contour_n2_$tag.json 2> gpurun_out/full_contour_n2_$tag.err; tail -c 1500 gpurun_out/full_contour_n2_$tag.json
