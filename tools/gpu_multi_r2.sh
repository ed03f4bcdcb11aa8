#!/bin/bash
# Round-2 multi-GPU evidence (one 8-GPU box): weak scaling of the path at 1/2/4/8 GPUs (BASELINE configs[1]), the bank-size
# sweep of configs[2] at 8 GPUs, the joint training step of configs[3] at 1 and 8 GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 600 python bench.py --steps 20 --warmup 5 --no-generator --no-extras 2>gpurun_out/r2_scale_n1.err | tail -1 > gpurun_out/r2_scale_n1.json
for n in ${SCALE_NS:-2 4 8}; do
  [ $n -le $N ] && timeout 600 bash -c "$(declare -f run); run $n $((29510+n)) bench.py --gpus $n --steps 20 --warmup 5" 2>gpurun_out/r2_scale_n$n.err | tail -1 > gpurun_out/r2_scale_n$n.json
done
for m in 512 1000 2000; do
  timeout 600 bash -c "$(declare -f run); run $N $((29540+m%7)) bench.py --gpus $N --steps 20 --warmup 5 --items $m" 2>gpurun_out/r2_items_${m}_n$N.err | tail -1 > gpurun_out/r2_items_${m}_n$N.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --items 2000 --no-generator --no-extras 2>gpurun_out/r2_items_2000_n1.err | tail -1 > gpurun_out/r2_items_2000_n1.json
timeout 600 python tools/train_step.py --steps 10 --warmup 3 --batch 8 2>&1 | tail -1 > gpurun_out/r2_train_n1.json
timeout 600 bash -c "$(declare -f run); run $N 29560 tools/train_step.py --steps 10 --warmup 3 --batch 8" 2>&1 | tail -1 > gpurun_out/r2_train_n$N.json
if [ -n "$SYNC_BN" ]; then
  timeout 600 bash -c "$(declare -f run); run $N 29561 tools/train_step.py --steps 10 --warmup 3 --batch 8 --sync-bn" 2>&1 | tail -1 > gpurun_out/r2_train_syncbn_n$N.json
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_scale_n*.json') + glob.glob('gpurun_out/r2_items_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value']), 'fps', round(d['ms_per_step'], 3), 'ms/step e2e', round(d['e2e']['value']), 'numa', d['e2e'].get('numa_node'))
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-300:])
for f in sorted(glob.glob('gpurun_out/r2_train_*.json')):
    print(f, open(f).read()[-400:])
PY
