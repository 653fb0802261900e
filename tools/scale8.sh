#!/bin/bash
# 8-GPU attribution runs (weak scaling, cfg2, 32 frames per GPU): default, without the all-gather, vote at 2 blocks/SM
o=gpurun_out
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 3 --no-matching --no-head-epilogue --no-cpu "$@" > $o/r02_scale8_$tag.json 2> $o/r02_scale8_$tag.err; python -c "
import json
d=json.load(open('$o/r02_scale8_$tag.json'))
print('$tag', 'fps=%.0f'%d['value'], 'ms=%.4f'%d['ms_per_step'], 'e2e', d['e2e'] and round(d['e2e']['value']), 'numa', d.get('numa'))"; }
python bench.py --steps 30 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_scale1_ref.json 2>/dev/null; python -c "
import json
d=json.load(open('$o/r02_scale1_ref.json')); print('1gpu fps=%.0f ms=%.4f e2e=%.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
run default
run nogather --no-gather --no-e2e
FPC_VOTE_BLOCKS_PER_SM=2 run bps2 --no-e2e
run nonuma --no-numa-bind
