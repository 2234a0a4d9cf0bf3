#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list of one bench command + full capture of the three hot kernels.
# Usage (under gpurun): bash profiles/run_profile.sh <tag> [size]
TAG=${1:-r01}
SIZE=${2:-8192}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --size $SIZE --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bfs_kernel|sweep_kernel|parse_kernel" -s 9 -c 3 \
    -f -o gpurun_out/prof_${TAG} python bench.py --size $SIZE --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
