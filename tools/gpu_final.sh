#!/bin/bash
cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph --no-streams > /dev/null 2>&1
timeout 600 python tools/generator_bench.py --batch 16 > gpurun_out/generator_bench_b16.txt 2>&1
timeout 600 python tools/bench_reductions.py > gpurun_out/reductions_bench.jsonl 2>&1
timeout 600 python tools/train_amft_bench.py > /dev/null 2>&1
cat gpurun_out/bench_n1.json | cut -c1-200
