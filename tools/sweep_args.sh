#!/bin/bash
# usage: ENVS="A=1" tools/sweep_args.sh "--pipeline-depth 4" "--pipeline-depth 8" ...
for a in "$@"; do
  env $ENVS python bench.py --steps 32 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu $a 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$ENVS $a:', 'step_ms=%.4f'%d['ms_per_step'], 'fps=%.0f'%d['value'])"
done
