#!/bin/bash
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | head -2 | cut -c1-200
timeout 600 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/train_amft_bench.py 2>&1 | tail -1 | cut -c1-200
