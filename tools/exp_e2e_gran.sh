#!/bin/bash
# e2e (zero-copy over PCIe) vs cudaLimitMaxL2FetchGranularity, and the gather address-stream micro-benchmark
o=gpurun_out
./build/microbench_gather > $o/r02_microbench_gather.txt 2>&1
for g in 0 32 64; do for d in 1 2; do
python bench.py --steps 10 --warmup 3 --no-matching --no-head-epilogue --no-cpu --l2-fetch-granularity $g --e2e-depth $d 2>$o/e2e_g$g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('gran=$g depth=$d', 'e2e fps=%.0f ms=%.3f'%(e['value'], e['ms_per_step']), 'link GB/s=%.1f achieved=%.1f'%(e['host_to_device_copy_gbs_all_ranks'], e['achieved_host_read_gbs_all_ranks']), 'step_ms=%.4f'%d['ms_per_step'], 'gather=%.4f'%d['kernel_ms']['k_gather'])" >> $o/r02_e2e_gran.txt 2>&1
done; done
cat $o/r02_microbench_gather.txt $o/r02_e2e_gran.txt; grep -h cudaLimit $o/e2e_g*.err
