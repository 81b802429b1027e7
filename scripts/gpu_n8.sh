#!/bin/bash
# 8-GPU cycle: the bench at N=8 (weak scaling, 1024 omega points) and the 20-shell full beta-decay contour at N=8.
tag=${1:-n8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 \
  > gpurun_out/bench_n8_$tag.json 2> gpurun_out/bench_n8_$tag.err; head -c 400 gpurun_out/bench_n8_$tag.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 1 \
  --full-contour --shells 20 > gpurun_out/full_contour_n8_$tag.json 2> gpurun_out/full_contour_n8_$tag.err; tail -c 1300 gpurun_out/full_contour_n8_$tag.json
