#!/bin/bash
# A/B cycle on the GPU box: GPU parity tests, then the default bench with and without an environment knob.
# Usage: scripts/gpu_ab.sh <tag> <ENVVAR=value for the B arm> [pytest -k expr]
tag=${1:-ab}; knob=$2; kexpr=$3
mkdir -p gpurun_out
if [ "$kexpr" != "none" ]; then
  if [ -n "$kexpr" ]; then sel=(-k "$kexpr"); else sel=(); fi
  (timeout 1500 python -m pytest tests -q -m gpu -x "${sel[@]}" 2>&1 | tail -25) > gpurun_out/pytest_gpu_$tag.log
  tail -5 gpurun_out/pytest_gpu_$tag.log
fi
for arm in A B; do
  if [ $arm = B ]; then [ -z "$knob" ] && break; export "$knob"; fi
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_$arm.json 2> gpurun_out/bench_${tag}_$arm.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_$arm.json"))
    print("$arm", "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_max_rel"], d["parity_points"], "conv", d["converged_fraction"])
except Exception as e:
    print("bench $arm failed", e)
PY
  tail -2 gpurun_out/bench_${tag}_$arm.err
done
