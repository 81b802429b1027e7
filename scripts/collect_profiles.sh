#!/bin/bash
# Copy the evidence of one scripts/gpu_evidence.sh run from gpurun_out/ (scratch) into profiles/ (tracked), derive the launch
# shares and the per-kernel ncu table, and write the SASS mnemonic census of the shipped library.
# Usage: scripts/collect_profiles.sh <gpurun tag> <profiles prefix>      e.g.  scripts/collect_profiles.sh r02b r02
tag=$1; pre=${2:-r02}
cd "$(dirname "$0")/.."
g=gpurun_out
cp $g/bench_default_$tag.json profiles/${pre}_bench_default.json
cp $g/bench_reference_$tag.json profiles/${pre}_bench_reference_arm.json
[ -f $g/pytest_gpu_$tag.log ] && cp $g/pytest_gpu_$tag.log profiles/${pre}_pytest_gpu.log
cp $g/smoke_$tag.log profiles/${pre}_smoke.log
for f in bench_2phase bench_12sh bench_20sh bench_24sh level0 fp64_peak; do [ -f $g/${f}_$tag.json ] && cp $g/${f}_$tag.json profiles/${pre}_$f.json; done
cp $g/smi_$tag.csv profiles/${pre}_smi.csv
cp $g/launches_$tag.csv profiles/${pre}_launches_16sh_128pts.csv
python scripts/launch_shares.py $g/launches_$tag.csv > profiles/${pre}_launch_shares.txt
python scripts/ncu_kernels.py $g/prof_${tag}_raw.csv profiles/${pre}_ncu_kernels \
  "python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline (8 omega points per launch, 16 shells)" > profiles/${pre}_ncu_kernels.txt
cp profiles/${pre}_ncu_kernels.json profiles/ncu_kernels.json
# SASS census: what the shipped library is made of (FP64 tensor path = DMMA; asynchronous copies = LDGSTS / UBLKCP)
python scripts/sass_census.py pynfam_b200/lib/libpnfam_b200.so > profiles/${pre}_sass_census.txt
tail -1 profiles/${pre}_sass_census.txt
