#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list of one bench command + full capture of the hot kernels.
# Usage (under gpurun): bash profiles/run_profile.sh <tag> [size] [solver]
TAG=${1:-r01}
SIZE=${2:-8192}
SOLVER=${3:-tiles}
mkdir -p gpurun_out
CMD="python bench.py --size $SIZE --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --solver $SOLVER"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    $CMD > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
if [ "$SOLVER" = "tiles" ]; then
  # one launch of each hot kernel of the 4th step (fused-parse path: 8 kernel launches per step, 2 synth kernels first)
  REGEX="tile_phase_a_kernel|tile_phase_c_kernel|slots_solve_kernel|slots_finalize_kernel|pit_scatter_kernel"; SKIP=15; COUNT=5
else
  REGEX="bfs_kernel|sweep_kernel|parse_kernel"; SKIP=9; COUNT=3
fi
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
