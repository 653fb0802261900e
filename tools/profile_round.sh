#!/bin/bash
# Regenerates the ncu evidence kept under profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/):
#   1. launch list of the bench command (eager launches so every kernel is its own ncu row)
#   2. `--set full` capture of the 15 path kernels of one step
#   3. `--set full` capture of the kernels of the "next" rows (matching, head-epilogue fusion)  [skipped with PATH_ONLY=1]
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu --no-matching --no-head-epilogue --pipeline-depth 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r01_launches.csv $BENCH > gpurun_out/r01_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 45 -c 15 -o gpurun_out/r01_path_full -f $BENCH > gpurun_out/r01_path_full.log 2>&1
if [ -z "$PATH_ONLY" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"^k_(argmax_runs_up4|gather|upsample|pack|mask_iou|match|paint)" -s 3 -c 11 -o gpurun_out/r01_extras_full -f python tools/extras_run.py > gpurun_out/r01_extras_full.log 2>&1
fi
tail -n 2 gpurun_out/r01_launches.log gpurun_out/r01_path_full.log | cut -c1-200
