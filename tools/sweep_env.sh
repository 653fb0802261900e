#!/bin/bash
# usage: tools/sweep_env.sh "A=1 B=2" "A=2 B=3" ...   (each argument = one environment for a short bench run)
for cfg in "$@"; do
  env $cfg python bench.py --steps 24 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
km=d['kernel_ms']
print('$cfg:', 'step_ms=%.4f'%d['ms_per_step'], 'fps=%.0f'%d['value'], 'argmax=%.4f gather=%.4f vote=%.4f settle=%.4f sum=%.3f'%(km['k_argmax_runs'],km['k_gather'],km['k_vote'],km['k_vote_settle'],sum(km.values())))"
done
