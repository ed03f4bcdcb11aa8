#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python tools/score_dataset.py --dataset ped2 --native --batch 64 2>&1 | tail -1 | tee gpurun_out/score_dataset_ped2_native_n1.json | cut -c1-600
timeout 900 python tools/score_dataset.py --dataset ped2 --native --batch 64 --graph 2>&1 | tail -1 | tee gpurun_out/score_dataset_ped2_native_graph_n1.json | cut -c1-600
