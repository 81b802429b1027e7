#!/bin/bash
# Evidence run (on the B200 box through gpurun): GPU parity tests, default bench (with CPU baseline), reference arm, ncu
# launch list of one bench step, one full ncu capture of every kernel family of the iteration, smoke.
# Usage: scripts/gpu_evidence.sh <tag> [skip-tests]
tag=${1:-ev}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  (timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu_$tag.log
  tail -3 gpurun_out/pytest_gpu_$tag.log
fi
timeout 900 python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
tail -c 2500 gpurun_out/bench_default_$tag.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
tail -c 700 gpurun_out/bench_reference_$tag.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_launch_$tag.log 2>&1
# (gpurun brings back at most 64 MiB: one iteration's worth of kernels, raw CSV exported on the box, big reports dropped)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sf2_|transform_|bro_|greens|fields|strength_kernel" -s 46 -c 23 \
  -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/b_ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/prof_$tag.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ]; then rm -f gpurun_out/prof_$tag.ncu-rep; fi
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke_$tag.log; cat gpurun_out/smoke_$tag.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv
