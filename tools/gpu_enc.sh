#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_memory.py tests/test_gpu_scoring.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-generator 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['breakdown'], d['roofline']['avg_launch_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"enc_tc" -s 4 -c 4 --csv --log-file gpurun_out/enc_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/enc_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
print([ (r[ik].split('(')[0][-24:], r[iv]) for r in rows[1:]])
PY
