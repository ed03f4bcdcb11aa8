#!/bin/bash
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r01_bench_n1_gen.json 2> gpurun_out/bench_err.log; tail -c 1500 gpurun_out/r01_bench_n1_gen.json
