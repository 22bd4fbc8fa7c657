python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py --deposit 2 --refill 8 --agg 24 --reps 2 --opts noabsorbed=0,with_abu=1 > gpurun_out/sweep_abu.log 2>&1
cut -c1-200 gpurun_out/sweep_abu.log | tail -3
