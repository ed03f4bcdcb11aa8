#!/bin/bash
# Profile evidence for profiles/: launch list of one eager step, full ncu captures of the three tcgen05 kernels, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair -s 12 -c 2 -f -o gpurun_out/prof_conv_pair \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"enc_tc_kernel|addr_tc_kernel|refine_kernel" -s 6 -c 3 -f -o gpurun_out/prof_mem \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
ls -la gpurun_out | tail -8
cat gpurun_out/bench_n1.json | cut -c1-300
