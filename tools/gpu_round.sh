#!/bin/bash
# One GPU-box session: full GPU test suite, the bench lines of every workload, the reference arm, and the ncu evidence.
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $o/r02_gputests.log
python bench.py --steps 20 --warmup 3 > $o/r02_bench_1gpu.json 2> $o/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/r02_bench_reference_arm.json 2> $o/r02_bench_reference_arm.err
python bench.py --workload cfg3 --steps 20 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_bench_cfg3_1gpu.json 2> $o/r02_bench_cfg3.err
python bench.py --workload cfg4 --steps 20 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_bench_cfg4_1gpu.json 2> $o/r02_bench_cfg4.err
python bench.py --workload cfg1 --steps 50 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_bench_cfg1_1gpu.json 2> $o/r02_bench_cfg1.err
python bench.py --workload cfg5 --steps 10 --warmup 3 > $o/r02_bench_cfg5_1gpu.json 2> $o/r02_bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 48 -c 64 --csv --log-file $o/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 64 -c 16 -o $o/r02_path_full python bench.py --steps 2 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu --no-graph --pipeline-depth 1 > /dev/null 2>&1
python bench.py --workload cfg1 --steps 100 --warmup 3 --no-matching --no-head-epilogue --no-cpu --no-e2e --pipeline-depth 1 > $o/r02_bench_cfg1_latency.json 2>/dev/null
python bench.py --steps 40 --warmup 3 --no-matching --no-head-epilogue --no-cpu --no-e2e --pipeline-depth 1 > $o/r02_bench_cfg2_latency.json 2>/dev/null
python tools/timeline.py --depth 4 > $o/r02_timeline_depth4.json 2>/dev/null
bash tools/sanitize.sh > $o/r02_sanitize_summary.txt 2>&1
tail -3 $o/r02_gputests.log
