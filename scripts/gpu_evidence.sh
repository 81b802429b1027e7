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
# same build, two-phase transform kernels (A/B of the fused transform)
PNFAM_B200_TRANSFORM_2PHASE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2phase_$tag.json 2> gpurun_out/bench_2phase_$tag.err
# the other basis sizes of BASELINE configs[4]
for sh in 12 20 24; do
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --shells $sh --points 64 > gpurun_out/bench_${sh}sh_$tag.json 2> gpurun_out/bench_${sh}sh_$tag.err
  python -c "
import json; d = json.load(open('gpurun_out/bench_${sh}sh_$tag.json')); print($sh, 'shells', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_max_rel'], d['parity_points'])"
done
# Level 0: one process per omega point
timeout 600 python scripts/level0.py > gpurun_out/level0_$tag.json 2> gpurun_out/level0_$tag.err; cat gpurun_out/level0_$tag.json
# FP64 peak probe with the clocks it ran at
python - <<PY > gpurun_out/fp64_peak_$tag.json
import json, subprocess, sys
sys.path.insert(0, '.')
from pynfam_b200 import gpu
vals = [gpu.dmma_peak_tflops(0) for _ in range(5)]
smi = subprocess.run(['nvidia-smi', '--query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_throttle_reasons.active', '--format=csv,noheader'],
                     capture_output=True, text=True).stdout.strip()
print(json.dumps({'probe': 'pnfam_b200_dmma_peak: register-resident mma.sync.m8n8k4.f64 loop on all SMs (csrc/cuda/solver.cu)', 'tflops': vals,
                  'theory': '148 SM x 128 flop/clk x 1.965 GHz = 37.2 TFLOP/s', 'nvidia_smi': smi}))
PY
cat gpurun_out/fp64_peak_$tag.json
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke_$tag.log; cat gpurun_out/smoke_$tag.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv
