#!/bin/bash
# k_gather with speculative class loads: parity, then grid / register-cap sweep, then the e2e depth sweep
o=gpurun_out
python -m pytest tests/test_fused_gpu.py tests/test_fullsize_parity_gpu.py tests/test_dropin_gpu.py tests/test_head_epilogue_gpu.py -m gpu -q -x 2>&1 | tail -3 > $o/r02_exp_gather.txt
tools/sweep_env.sh "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=32" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=16" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=8" "FPC_GATHER_MINB=4 FPC_GATHER_BLOCKS_PER_SM=4" \
  "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=32" "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=12" "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=6" "FPC_GATHER_MINB=3 FPC_GATHER_BLOCKS_PER_SM=3" >> $o/r02_exp_gather.txt 2>&1
for d in 1 2 3 4; do
python bench.py --steps 10 --warmup 3 --no-matching --no-head-epilogue --no-cpu --e2e-depth $d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('e2e depth=$d', 'fps=%.0f ms=%.3f'%(e['value'], e['ms_per_step']), 'link GB/s=%.1f achieved=%.1f'%(e['host_to_device_copy_gbs_all_ranks'], e['achieved_host_read_gbs_all_ranks']))" >> $o/r02_exp_gather.txt 2>&1
done
cat $o/r02_exp_gather.txt
