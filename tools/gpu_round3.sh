#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_memory.py -q -m gpu --timeout 600 -k "tensor" 2>&1 | tail -40 | tee gpurun_out/pytest_tensor.log
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_addressing.py --quick --out gpurun_out/addressing_quick.json 2>&1 | tail -12 | tee gpurun_out/addressing_quick.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_n1.json
