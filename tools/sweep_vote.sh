#!/bin/bash
# scratch: A/B of the vote kernel's work-item size, tail divisor and residency (run on the GPU box)
out=gpurun_out/r02_sweep_vote.txt
: > $out
for px in ${PXS:-1024}; do for bps in ${BPS:-2 3}; do for td in ${TDS:-1 4}; do
  FPC_VOTE_ITEM_PX=$px FPC_VOTE_BLOCKS_PER_SM=$bps FPC_VOTE_TAIL_DIV=$td python bench.py --steps 20 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('px=$px bps=$bps tail=$td', 'step_ms=%.4f'%d['ms_per_step'], 'vote_ms=%.4f'%d['kernel_ms']['k_vote'], 'frac=%.3f'%d['roofline_fp32_voting']['frac'], 'fps=%.0f'%d['value'])" >> $out
done; done; done
cat $out
