#!/bin/bash
# Last sanity pass of a round: every GPU test, smoke(), the default bench, a training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
timeout 600 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -1 > gpurun_out/train_n1.json
cut -c1-200 gpurun_out/bench_n1.json; cut -c1-330 gpurun_out/train_n1.json
