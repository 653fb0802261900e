#!/bin/bash
for pad in 0 27000 36000 56000 75000 110000 200000; do
  FPC_ARGMAX_PAD_SMEM=$pad python bench.py --steps 10 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pad=$pad', 'argmax_ms=%.4f'%d['kernel_ms']['k_argmax_runs'], 'step_ms=%.4f'%d['ms_per_step'])"
done
