#!/bin/bash
# Round-2 evidence run A (one GPU): full GPU test suite, smoke, bench, launch list, full ncu capture of the q conv kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-extras --no-graph --no-streams > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair_kernel -s 8 -c 4 -f -o gpurun_out/r2_prof_conv_q \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-extras --no-graph --no-streams > /dev/null 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_smoke.log; tail -c 1500 gpurun_out/r2_bench_n1.err; cut -c1-3000 gpurun_out/r2_bench_n1.json
