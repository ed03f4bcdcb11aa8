#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --durations=12 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
