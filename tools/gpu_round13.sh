#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 python tools/train_amft_bench.py 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_memory.py tests/test_gpu_amft.py -q -m gpu --timeout 1400 \
   -k "golden or tensor_dec or tensor_enc or conv1x1_engine or pair_kernel or training_forward or no_residual or k3" 2>&1 | grep -vE "^\s*$" | tail -15 | tee gpurun_out/sanitizer_memcheck.log
