#!/bin/bash
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | tail -7
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400
