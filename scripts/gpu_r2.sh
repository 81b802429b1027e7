#!/bin/bash
# Round-2 GPU cycle (run through gpurun): GPU parity tests, parity probes against the production-size fixtures, bench
# with a sweep over the slot count, launch list.   Usage: scripts/gpu_r2.sh <tag> [slots list] [pytest -k expr]
tag=${1:-r2}
slots=${2:-"0"}
kexpr=${3:-""}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  (timeout 1500 python -m pytest tests -q -m gpu -x -k "$kexpr" 2>&1 | tail -15) > gpurun_out/pytest_gpu_$tag.log
else
  (timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15) > gpurun_out/pytest_gpu_$tag.log
fi
cat gpurun_out/pytest_gpu_$tag.log
for s in $slots; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --slots $s > gpurun_out/bench_${tag}_s$s.json 2> gpurun_out/bench_${tag}_s$s.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_s$s.json"))
    print("slots", "$s", d["problem"], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_max_rel"], d["parity_points"],
          "conv", d["converged_fraction"], "frac", d["roofline"]["frac"], d["roofline"]["density"], d["roofline"]["projection"])
except Exception as e:
    print("bench failed", e)
PY
  tail -2 gpurun_out/bench_${tag}_s$s.err
done
