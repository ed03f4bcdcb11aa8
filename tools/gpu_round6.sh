#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
rm -f gpurun_out/wgrad_probe.log
for c in w64 w64b w512 w512p3; do
  timeout 120 python tools/conv_probe.py $c 2>&1 | tail -8 | tee -a gpurun_out/wgrad_probe.log
done
timeout 900 python -m pytest tests/test_gpu_amft.py -q -m gpu --timeout 600 2>&1 | tail -30 | tee gpurun_out/pytest_amft.log
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
