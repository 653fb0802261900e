#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests (racecheck is ~100x slower than native: small scenes only)
o=gpurun_out
sel="tests/test_dropin_gpu.py::test_ransac_voting_layer_v3_bit_exact_votes tests/test_fused_gpu.py::test_fused_vs_oracle tests/test_dropin_gpu.py::test_voting_code_paths_bit_exact"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest $sel -m gpu -q -x -k "three_frames or touching or 129 or 1030" > $o/r02_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_fused_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -q -x -k "not larger_than and not cfg4" > $o/r02_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool synccheck python -m pytest tests/test_fused_gpu.py::test_fused_vs_oracle -m gpu -q -x -k "three_frames" > $o/r02_synccheck.log 2>&1
for f in racecheck memcheck synccheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" $o/r02_$f.log | tail -5; done
