#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m ammcnet_aaai2021_b200.build 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_memory.py tests/test_gpu_amft.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/bench_reductions.py 2>&1 | tail -1 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gz_kernel|gx_kernel|genc_w|gdec|conv_igemm_pair|pack_planes|pack_nhwc64|channel_sum|read_planes|conv_wgrad|bank_transpose" -s 44 -c 12 --csv --log-file gpurun_out/mem_bwd_launches.csv \
    python tools/bench_reductions.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/mem_bwd_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]: print(r[ik].split('(')[0][:44], r[iv])
PY
