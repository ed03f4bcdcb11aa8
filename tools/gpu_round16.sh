#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -12
timeout 300 python tools/preprocess_bench.py 2>&1 | tee gpurun_out/preprocess_bench.jsonl | cut -c1-220
