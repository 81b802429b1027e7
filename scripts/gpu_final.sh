#!/bin/bash
# Round-end evidence run (on the B200 box through gpurun): GPU parity tests, default bench (with CPU baseline),
# reference arm, ncu launch list and a full ncu capture of the two dominant kernels.  Usage: scripts/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu_$tag.log
cat gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
tail -c 1500 gpurun_out/bench_default_$tag.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
tail -c 600 gpurun_out/bench_reference_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_launch_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"density_kernel|projection_kernel" -s 8 -c 6 -o gpurun_out/prof_$tag \
  python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/b_ncu_$tag.log 2>&1
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke_$tag.log; cat gpurun_out/smoke_$tag.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv
