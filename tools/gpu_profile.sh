#!/bin/bash
# One-GPU evidence run for profiles/: tests, smoke, bench (+reference arm), launch list, full ncu captures of the tcgen05
# kernels, generator / reductions / preprocessing / training benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph --no-streams > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair -s 14 -c 2 -f -o gpurun_out/prof_conv_pair \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph --no-streams > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"enc_tc_kernel|addr_tc_kernel|refine_kernel" -s 6 -c 3 -f -o gpurun_out/prof_mem \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph --no-streams > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair -s 12 -c 1 -f -o gpurun_out/prof_dec_pair \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-graph --no-streams > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -f -o gpurun_out/prof_halo64 \
    python tools/layer_once.py 64 64 256 256 16 1 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -f -o gpurun_out/prof_halo128 \
    python tools/layer_once.py 256 128 128 128 16 1 3 > /dev/null 2>&1
timeout 600 python tools/generator_bench.py --batch 16 > gpurun_out/generator_bench_b16.txt 2>&1
timeout 600 python tools/generator_bench.py --batch 64 --steps 5 > gpurun_out/generator_bench_b64.txt 2>&1
timeout 600 python tools/generator_layers.py --batch 16 > gpurun_out/generator_layers_b16.txt 2>&1
timeout 600 python tools/bench_reductions.py > gpurun_out/reductions_bench.jsonl 2>&1
timeout 300 python tools/preprocess_bench.py > gpurun_out/preprocess_bench.jsonl 2>&1
timeout 900 python tools/score_dataset.py --dataset ped2 --native --batch 64 2>&1 | tail -1 > gpurun_out/score_dataset_ped2_native_n1.json
timeout 600 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -1 > gpurun_out/train_n1.json
timeout 600 python tools/train_amft_bench.py 2>&1 | tail -4 > gpurun_out/train_amft_bench.txt
ls gpurun_out | tr '\n' ' '
cat gpurun_out/bench_n1.json | cut -c1-260
