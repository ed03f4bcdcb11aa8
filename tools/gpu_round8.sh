#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    a=agg.setdefault(row['Kernel Name'][:56],[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"{k:58s} n={n:4d} avg={t/n:8.1f} us share={100*t/tot:5.1f}%")
PY
