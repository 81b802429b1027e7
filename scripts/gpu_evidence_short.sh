#!/bin/bash
# Reduced evidence run (when little GPU time is left): default bench with CPU baseline, launch list, one full ncu capture of
# every kernel family, the two-phase A/B and the 12- / 24-shell benches.  Usage: scripts/gpu_evidence_short.sh <tag>
tag=${1:-ev}
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
tail -c 600 gpurun_out/bench_default_$tag.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_launch_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"sf2_|transform_|bro_|greens|fields|strength_kernel" -s 46 -c 23 \
  -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/b_ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
PNFAM_B200_TRANSFORM_2PHASE=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2phase_$tag.json 2> gpurun_out/bench_2phase_$tag.err
for sh in 12 24; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --shells $sh --points 64 > gpurun_out/bench_${sh}sh_$tag.json 2> gpurun_out/bench_${sh}sh_$tag.err
done
echo done
