#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md), run under gpurun on ONE GPU:
#   bash profiles/run_profile_r02.sh <tag> [size]
# 1. launch list of one bench command (gpu__time_duration.sum, --clock-control none): kernel SHARES of the step
# 2. ncu --set full of one launch of every hot kernel of the headline step (phase A / solve / finalize / phase C)
# 3. ncu --set full of the tile-dataflow sweeps (Strahler, float32 accuflux, HAND) on the same raster
TAG=${1:-r02}
SIZE=${2:-8192}
mkdir -p gpurun_out
CMD="python bench.py --size $SIZE --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    $CMD > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"tile_phase_a_kernel|tile_phase_c_kernel|slots_solve_kernel|slots_finalize_kernel|pit_scatter_kernel" -s 15 -c 5 \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_up_sweep_kernel|tile_down_sweep_kernel" -c 3 \
    -f -o gpurun_out/prof_${TAG}_sweeps python profiles/scripts/sweep_case.py $SIZE 1 > gpurun_out/ncu_full_${TAG}_sweeps.log 2>&1
ls -la gpurun_out/ | tail -6
