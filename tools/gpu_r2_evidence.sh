#!/bin/bash
# Round-2 evidence run (one GPU): everything that ends up under profiles/r02_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-extras --no-graph --no-streams"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2e_bench_n1.err | tail -1 > gpurun_out/r2e_bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 2>&1 | tail -1 > gpurun_out/r2e_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2e_launches.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair_kernel -s 4 -c 6 -f -o gpurun_out/r2e_prof_pair $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mem_front_kernel -s 2 -c 2 -f -o gpurun_out/r2e_prof_front $B > /dev/null 2>&1
timeout 300 python tools/convq_probe.py 64 > gpurun_out/r2e_convq_probe_b64.json 2>&1
timeout 400 python tools/bench_addressing.py --quick > gpurun_out/r2e_addressing_quick.txt 2>&1
timeout 300 python tools/bench_reductions.py > gpurun_out/r2e_reductions_bench.jsonl 2>&1
timeout 400 python tools/train_amft_bench.py 2>&1 | tail -60 > gpurun_out/r2e_train_amft_bench.txt
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2e_sanitizer.log \
    python -m pytest tests/test_gpu_memory.py tests/test_gpu_amft.py -m gpu -x -q \
    -k "fused_front or q_conv or golden or prepared or staged" 2>&1 | tail -3 > gpurun_out/r2e_sanitizer_pytest.log
tail -3 gpurun_out/r2e_pytest_gpu.log; tail -2 gpurun_out/r2e_smoke.log; tail -c 300 gpurun_out/r2e_bench_n1.err; cut -c1-400 gpurun_out/r2e_bench_n1.json
tail -3 gpurun_out/r2e_sanitizer_pytest.log; tail -3 gpurun_out/r2e_sanitizer.log
