#!/bin/bash
# usage: VAR=name VALS="0 1" BPS="2 3" tools/sweep_ab.sh
for v in $VALS; do for bps in ${BPS:-3}; do
  env $VAR=$v FPC_VOTE_BLOCKS_PER_SM=$bps python bench.py --steps 24 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
km=d['kernel_ms']
print('$VAR=$v bps=$bps', 'step_ms=%.4f'%d['ms_per_step'], 'fps=%.0f'%d['value'], 'argmax=%.4f gather=%.4f vote=%.4f settle=%.4f'%(km['k_argmax_runs'],km['k_gather'],km['k_vote'],km['k_vote_settle']))"
done; done
