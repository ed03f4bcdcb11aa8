#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-generator 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['breakdown'], d['roofline']['avg_launch_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"enc_tc|conv_igemm_pair" -s 8 -c 8 --csv --log-file gpurun_out/enc_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph > /dev/null 2>&1
grep -E "enc_tc|conv_igemm" gpurun_out/enc_launches.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | cut -c1-60
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | head -2 | cut -c1-200
