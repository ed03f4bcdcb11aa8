#!/bin/bash
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 python tools/generator_layers.py --batch 16 2>&1 | tail -24
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | tail -6
