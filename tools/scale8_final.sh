#!/bin/bash
# 8-GPU box: cfg2 (the driver's SCALE shape) at 8 GPUs with e2e + per-rank times, cfg5 at 2 / 4 / 8 GPUs
o=gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
tr 8 --steps 30 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_bench_8gpu.json 2> $o/r02_bench_8gpu.err
python bench.py --steps 30 --warmup 3 --no-matching --no-head-epilogue --no-cpu > $o/r02_bench_1gpu_samebox.json 2>/dev/null
for n in 8 4 2; do
  tr $n --workload cfg5 --steps 10 --warmup 3 > $o/r02_bench_cfg5_${n}gpu.json 2> $o/r02_bench_cfg5_${n}gpu.err
done
python - <<'PY'
import json
for f in ["r02_bench_1gpu_samebox","r02_bench_8gpu","r02_bench_cfg5_2gpu","r02_bench_cfg5_4gpu","r02_bench_cfg5_8gpu"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, "fps=%.0f ms=%.4f"%(d["value"], d["ms_per_step"]), "e2e=%s"%(round(e["value"]) if e else None), e.get("host_to_device_copy_gbs_all_ranks"), d.get("ms_per_step_per_rank"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
