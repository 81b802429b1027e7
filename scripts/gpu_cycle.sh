#!/bin/bash
# One GPU iteration cycle (run on the B200 box through gpurun): GPU parity tests, a short bench, and an ncu capture
# of the dominant kernels.  Usage: scripts/gpu_cycle.sh <tag> [ncu_kernel_regex]
tag=${1:-x}
rx=${2:-sf_density_kernel|sf_projection_kernel}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15) > gpurun_out/pytest_gpu_$tag.log
cat gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"])
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/bench_$tag.err
if [ "$rx" != "none" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 8 -c 6 -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/b_ncu_$tag.log 2>&1
fi
