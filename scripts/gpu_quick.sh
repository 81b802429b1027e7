#!/bin/bash
# Quick GPU check (through gpurun): a few parity tests, a short bench and the ncu launch list of one bench step.
# Usage: scripts/gpu_quick.sh <tag> [pytest -k expression]
tag=${1:-q}
kexpr=${2:-"trajectory or whole_contour or calc_hamiltonian"}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -x -k "$kexpr" 2>&1 | tail -5) > gpurun_out/pytest_gpu_$tag.log
cat gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print(d["value"], d["e2e"]["value"], d["ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_launch_$tag.log 2>&1
python scripts/launch_shares.py gpurun_out/launches_$tag.csv | head -14
