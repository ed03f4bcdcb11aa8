#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/score_dataset.py --dataset ped2 --size 128 2>&1 | tail -2 | tee gpurun_out/score_ped2_n$N.json
for m in 512 1000 2000; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --items $m 2>&1 | tail -1 > gpurun_out/bench_items_$m.json
done
python - <<'PY'
import json
for m in (512, 1000, 2000):
    d = json.loads(open('gpurun_out/bench_items_%d.json' % m).read().strip().splitlines()[-1])
    print('M', m, round(d['value']), 'fps', round(d['ms_per_step'], 3), 'ms/step', d['breakdown'])
PY
