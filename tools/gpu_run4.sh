# look-ahead kernel: GPU tests, then the two bench launches for SOC_AHEAD = 0 (lean), 1 (3 CTAs/SM), 2 (4 CTAs/SM)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for a in 0 1 2; do
  echo "SOC_AHEAD=$a" >> gpurun_out/sweep_ahead.log
  SOC_AHEAD=$a timeout 300 python tools/sweep.py --deposit 0,2 --refill 8 --agg 24 --reps 3 >> gpurun_out/sweep_ahead.log 2>&1
done
cat gpurun_out/sweep_ahead.log
