#!/bin/bash
# Final round-2 evidence run (one GPU): everything that ends up under profiles/r02_* after the second half of the round
# (addressing tail, bf16 feature I/O, training weight gradient).  Same structure as gpu_r2_evidence.sh, prefix r2f_.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-generator --no-extras --no-graph --no-streams"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2f_bench_n1.err | tail -1 > gpurun_out/r2f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 2>&1 | tail -1 > gpurun_out/r2f_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_pair_kernel -s 4 -c 6 -f -o gpurun_out/r2f_prof_pair $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mem_front_kernel -s 2 -c 2 -f -o gpurun_out/r2f_prof_front $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:addr_tc_kernel|addr_tail_kernel" -c 2 -f -o gpurun_out/r2f_prof_addr python tools/addr_once.py 65536 2048 512 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv_wgrad -s 2 -c 1 -f -o gpurun_out/r2f_prof_wgrad python tools/amft_train_once.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2f_amft_train_launches.csv python tools/amft_train_once.py > /dev/null 2>&1
timeout 300 python tools/convq_probe.py 64 > gpurun_out/r2f_convq_probe_b64.json 2>&1
timeout 400 python tools/bench_addressing.py --quick > gpurun_out/r2f_addressing_quick.txt 2>&1
timeout 400 python tools/bench_addressing.py --points "65536,2048,512;262144,8192,1024;1048576,2048,256;16384,8192,1024;4096,1024,512" > gpurun_out/r2f_addressing_points.txt 2>&1
timeout 300 python tools/bench_reductions.py > gpurun_out/r2f_reductions_bench.jsonl 2>&1
timeout 400 python tools/train_amft_bench.py 2>&1 | tail -60 > gpurun_out/r2f_train_amft_bench.txt
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2f_sanitizer.log \
    python -m pytest tests/test_gpu_memory.py tests/test_gpu_bf16_io.py tests/test_gpu_amft.py -m gpu -x -q \
    -k "tensor_path or filter_decides or ema_statistics or bf16_io or fused_front or gradients" 2>&1 | tail -3 > gpurun_out/r2f_sanitizer_pytest.log
tail -3 gpurun_out/r2f_pytest_gpu.log; tail -2 gpurun_out/r2f_smoke.log; tail -c 300 gpurun_out/r2f_bench_n1.err; cut -c1-300 gpurun_out/r2f_bench_n1.json
tail -3 gpurun_out/r2f_sanitizer_pytest.log; tail -3 gpurun_out/r2f_sanitizer.log
