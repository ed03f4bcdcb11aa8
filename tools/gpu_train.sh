#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -3 | tee gpurun_out/train_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -3 | tee gpurun_out/train_n$N.json
