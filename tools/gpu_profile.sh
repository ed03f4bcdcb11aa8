#!/bin/bash
# Profile evidence for profiles/: tests, smoke, bench (+reference arm), launch lists, full ncu captures of the tcgen05 kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair -s 14 -c 2 -f -o gpurun_out/prof_conv_pair \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -f -o gpurun_out/prof_halo64 \
    python tools/layer_once.py 64 64 256 256 16 1 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -f -o gpurun_out/prof_halo128 \
    python tools/layer_once.py 256 128 128 128 16 1 3 > /dev/null 2>&1
timeout 600 python tools/generator_bench.py --batch 16 > gpurun_out/generator_bench_b16.txt 2>&1
timeout 600 python tools/generator_bench.py --batch 64 --steps 5 > gpurun_out/generator_bench_b64.txt 2>&1
tail -3 gpurun_out/generator_bench_b64.txt | cut -c1-250
ls -la gpurun_out | tail -12
cat gpurun_out/bench_n1.json | cut -c1-300
