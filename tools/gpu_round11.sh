#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1.json
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --pair-unfused 2>&1 | tail -1 > gpurun_out/bench_unfused.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair -s 14 -c 2 -f -o gpurun_out/prof_conv_pair \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
