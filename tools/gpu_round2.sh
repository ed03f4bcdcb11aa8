#!/bin/bash
# Full GPU pass: tests, smoke, bench (both arms), ncu launch list + one full capture of the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 16 -c 2 -f -o gpurun_out/prof_conv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_stdout.log 2>&1
ls -la gpurun_out
