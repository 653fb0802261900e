#!/bin/bash
# warp instructions and duration of every path kernel (one batch of the default workload), via ncu with two metrics only
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'^k_|::k_' -s 64 -c 16 --csv --log-file gpurun_out/r02_inst_counts.csv python bench.py --steps 2 --warmup 3 --no-matching --no-head-epilogue --no-e2e --no-cpu --no-graph --pipeline-depth 1 "$@" > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_inst_counts.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value')
d={}
for r in rows[1:]:
    name=r[ik].split('(')[0].replace('void ','').replace('fpc::','').split('<')[0]
    d.setdefault(name,{})[r[im]]=float(r[iv].replace(',',''))
tot=0
for k,v in d.items():
    print(f"{k:22s} {v.get('smsp__inst_executed.sum',0)/1e6:8.2f} M warp-instr  {v.get('gpu__time_duration.sum',0)/1e3:8.1f} us")
    tot+=v.get('smsp__inst_executed.sum',0)
print(f"{'total':22s} {tot/1e6:8.2f} M")
PY
