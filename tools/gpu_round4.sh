#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/addr_launches.csv python tools/addr_once.py > gpurun_out/addr_once.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/addr_launches.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        print("%-70s %10s %s" % (row['Kernel Name'][:70], row['Metric Value'], row['Metric Unit']))
PY
