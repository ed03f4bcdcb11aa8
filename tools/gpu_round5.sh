#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_addressing.py --quick --out gpurun_out/addressing_quick.json 2>&1 | tail -8 | tee gpurun_out/addressing_quick.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/addr_launches.csv python tools/addr_once.py > gpurun_out/addr_once.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/addr_launches.csv') if not l.startswith('==')]
seen=set()
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        print("%-60s %10s %s" % (row['Kernel Name'][:60], row['Metric Value'], row['Metric Unit']))
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_n1.json
