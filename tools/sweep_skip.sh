#!/bin/bash
out=gpurun_out/r02_sweep_skip.txt
: > $out
for sk in 0 1 2 3 8 11 4 15; do
  FPC_VOTE_DEBUG_SKIP=$sk FPC_VOTE_ITEM_PX=1024 FPC_VOTE_BLOCKS_PER_SM=3 FPC_VOTE_TAIL_DIV=1 python bench.py --steps 20 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('skip=$sk', 'step_ms=%.4f'%d['ms_per_step'], 'vote_ms=%.4f'%d['kernel_ms']['k_vote'])" >> $out
done
cat $out
