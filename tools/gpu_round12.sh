#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/score_dataset.py --dataset ped2 --size 256 2>&1 | tail -1 | tee gpurun_out/score_ped2_256_eager.json
timeout 600 python tools/score_dataset.py --dataset ped2 --size 256 --graph 2>&1 | tail -1 | tee gpurun_out/score_ped2_256_graph.json
timeout 1200 python tools/bench_addressing.py --out gpurun_out/addressing_sweep.json 2>&1 | tail -75 > gpurun_out/addressing_sweep.log
tail -72 gpurun_out/addressing_sweep.log
