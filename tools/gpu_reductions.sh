#!/bin/bash
# North-star kernels (b) and (c): roofline microbench + per-kernel times of the memory-module backward.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 python tools/bench_reductions.py 2>&1 | tee gpurun_out/reductions_bench.jsonl | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"gz_kernel|gx_kernel|genc_w|gdec|conv_igemm_pair|pack_planes|pack_nhwc64|channel_sum|read_planes|conv_wgrad|bank_transpose" \
    -s 44 -c 12 --csv --log-file gpurun_out/mem_bwd_launches.csv python tools/bench_reductions.py > /dev/null 2>&1
