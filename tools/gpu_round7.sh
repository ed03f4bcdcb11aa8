#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 120 python tools/conv_probe.py c512 c512p3 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_amft.py -q -m gpu --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_amft.log
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pair.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-pair 2>&1 | tail -1 | tee gpurun_out/bench_nopair.json
