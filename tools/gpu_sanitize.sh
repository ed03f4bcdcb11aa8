#!/bin/bash
# compute-sanitizer memcheck over the kernels added after the first sanitizer pass (halo conv, U-Net data movement,
# preprocessing, losses, reworked enc / backward); bounded so a slow box cannot run away with the budget
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build > /dev/null 2>&1
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_new_kernels.log \
    python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_losses.py tests/test_gpu_generator.py -m gpu -x -q \
    -k "preprocess or losses or halo or transposed or maxpool or padded or windows or wider" 2>&1 | tail -3
echo "exit=$?"
tail -4 gpurun_out/sanitizer_new_kernels.log
