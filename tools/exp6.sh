#!/bin/bash
# A/B: last-block-done folded scans (12 launches) vs separate single-block scan kernels (16 launches), same scan code, 8-pixel notes
o=gpurun_out
cp fastposecnn_b200/libfpc_b200.so build/libfpc_keep.so
: > $o/r02_exp6.txt
for rep in 1 2 3; do for lib in nofold fold; do
  cp build/libfpc_$lib.so fastposecnn_b200/libfpc_b200.so
  for args in "--workload cfg2 --pipeline-depth 4" "--workload cfg2 --pipeline-depth 1" "--workload cfg1 --pipeline-depth 1 --steps 100"; do
  python bench.py --steps 40 --warmup 3 --no-matching --no-head-epilogue --no-cpu --no-e2e $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib rep$rep $args: ms=%.4f'%d['ms_per_step'], 'sum=%.4f'%sum(d['kernel_ms'].values()))" >> $o/r02_exp6.txt 2>&1
  done
done; done
cp build/libfpc_keep.so fastposecnn_b200/libfpc_b200.so
sort $o/r02_exp6.txt
