#!/bin/bash
out=gpurun_out/r02_sweep_depth.txt
: > $out
for bps in 3 2; do for depth in 2 4 6 8 12; do
  FPC_VOTE_BLOCKS_PER_SM=$bps python bench.py --steps 24 --warmup 3 --pipeline-depth $depth --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bps=$bps depth=$depth', 'step_ms=%.4f'%d['ms_per_step'], 'fps=%.0f'%d['value'])" >> $out
done; done
cat $out
