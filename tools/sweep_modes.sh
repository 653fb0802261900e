#!/bin/bash
run() { python bench.py --steps 24 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'step_ms=%.4f'%d['ms_per_step'], 'fps=%.0f'%d['value'], 'sum_kernels=%.4f'%sum(d['kernel_ms'].values()))"; }
for bps in 3 1; do
export FPC_VOTE_BLOCKS_PER_SM=$bps
echo "bps=$bps"
run --pipeline-depth 4
run --pipeline-depth 4 --no-graph
run --pipeline-depth 4 --single-stream
run --pipeline-depth 1
run --pipeline-depth 1 --no-graph
done
