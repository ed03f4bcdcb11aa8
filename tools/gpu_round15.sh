#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 290 -c 80 --csv --log-file gpurun_out/generator_launches.csv \
    python tools/generator_layers.py --batch 16 > /dev/null 2>&1
wc -l gpurun_out/generator_launches.csv
