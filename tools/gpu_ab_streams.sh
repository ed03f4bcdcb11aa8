#!/bin/bash
# A/B: memory modules + PSNR on three streams vs one (bench.py), generator engine with per-stream concurrency
cd "$(dirname "$0")/.."
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_gpu_scoring.py -m gpu -x -q 2>&1 | tail -3
for f in "" "--no-streams"; do
  timeout 600 python bench.py --no-cpu-baseline --no-generator $f 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value']), d['ms_per_step'], d['breakdown'])"
done
timeout 600 python tools/generator_bench.py --batch 16 2>&1 | head -2 | cut -c1-200
