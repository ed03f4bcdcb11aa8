#!/bin/bash
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_amft.py tests/test_gpu_memory.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | head -2 | cut -c1-200
timeout 300 python tools/layer_once.py 64 64 256 256 16 1 3
timeout 300 python tools/layer_once.py 64 64 256 256 16 1 1
