#!/bin/bash
# Multi-GPU: bench under torchrun at N = all visible GPUs, plus N=1 on the same box for the ratio.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_multi_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510+n)) \
        bench.py --gpus $n --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_multi_n$n.json
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_multi_ref_n$N.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_multi_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value']), 'fps', round(d['ms_per_step'], 3), 'ms/step e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
