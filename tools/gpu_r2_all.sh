#!/bin/bash
(time python -m pytest tests -m gpu -q --durations=12) > gpurun_out/r2_gpu_all_b.log 2>&1
tail -25 gpurun_out/r2_gpu_all_b.log
echo "== 512 diag"
SOC_DOMAIN_VERBOSE=1 python - <<'PY' > gpurun_out/r2_diag512.log 2>&1
import sys, time
sys.path.insert(0, '.')
import bench, json
from soc_b200 import backend
import torch
peak = 6539.2
for i in range(2):
    t = time.time()
    line = bench.extra_grid(backend, 0, 0, 1, 512, peak, steps=2, cpu_seconds=0)
    print(json.dumps(line), time.time() - t)
PY
tail -12 gpurun_out/r2_diag512.log
