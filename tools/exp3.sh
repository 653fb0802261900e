#!/bin/bash
# folded scans + k_gather0 + 8-pixel notes: full GPU suite, then A/B sweeps (gather kernel / register cap / grid; note size)
o=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $o/r02_exp3.txt
tools/sweep_env.sh "FPC_GATHER_GENERIC=1" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=32" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=8" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=4" \
  "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=32" "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=6" "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=3" >> $o/r02_exp3.txt 2>&1
cp fastposecnn_b200/libfpc_b200.so build/libfpc_note8.so
for v in 16 4 8; do
  cp build/libfpc_note$v.so fastposecnn_b200/libfpc_b200.so
  echo "note_px=$v" >> $o/r02_exp3.txt
  tools/sweep_env.sh "FPC_VOTE_BLOCKS_PER_SM=2" "FPC_VOTE_BLOCKS_PER_SM=3" >> $o/r02_exp3.txt 2>&1
done
python bench.py --workload cfg1 --steps 50 --warmup 3 --no-matching --no-head-epilogue --no-cpu --no-e2e --pipeline-depth 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg1 depth1 ms=%.4f'%d['ms_per_step'], d['kernel_ms'])" >> $o/r02_exp3.txt 2>&1
python bench.py --steps 30 --warmup 3 --no-matching --no-head-epilogue --no-cpu --no-e2e --pipeline-depth 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 depth1 ms=%.4f'%d['ms_per_step'], d['kernel_ms'])" >> $o/r02_exp3.txt 2>&1
cat $o/r02_exp3.txt
