#!/bin/bash
# First GPU pass: per-file pytest (separate processes so one trap cannot take the others down) + conv probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
for f in memory scoring; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_$f.log
done
for c in g64 g64b g128 g256 g512 c64 c64s c512 c512p3; do
  timeout 120 python tools/conv_probe.py $c 2>&1 | tail -25 | tee -a gpurun_out/conv_probe.log
done
timeout 900 python -m pytest tests/test_gpu_amft.py -q -m gpu --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_amft.log
