#!/bin/bash
# Round-2 last evidence run (one GPU): all GPU tests with the training objectives, smoke(), the default bench, the objectives
# micro-benchmark, one adversarial training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2g_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
timeout 200 python tools/losses_bench.py > gpurun_out/r2g_losses_bench.jsonl 2>&1
timeout 300 python tools/train_step.py --gan --flownet --steps 10 --warmup 3 --batch 8 2>&1 | tail -2 > gpurun_out/r2g_train_gan_flownet_n1.json
timeout 300 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -1 > gpurun_out/r2g_train_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_objective_launches.csv python tools/objective_once.py > /dev/null 2>&1
timeout 100 python tools/bench_reductions.py 2>&1 | grep '^{' > gpurun_out/r2g_reductions_bench.jsonl
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_mem_bwd_launches.csv python tools/mem_bwd_once.py > /dev/null 2>&1
timeout 600 python bench.py 2>gpurun_out/r2g_bench_n1.err | tail -1 > gpurun_out/r2g_bench_n1.json
tail -6 gpurun_out/r2g_pytest_gpu.log; tail -2 gpurun_out/r2g_smoke.log; tail -4 gpurun_out/r2g_losses_bench.jsonl | cut -c1-400
cut -c1-400 gpurun_out/r2g_train_gan_flownet_n1.json; cut -c1-300 gpurun_out/r2g_train_n1.json; tail -1 gpurun_out/r2g_reductions_bench.jsonl | cut -c1-200; tail -c 300 gpurun_out/r2g_bench_n1.err; cut -c1-250 gpurun_out/r2g_bench_n1.json
