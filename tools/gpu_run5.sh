python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for a in 1 3 2; do
  echo "SOC_AHEAD=$a" >> gpurun_out/sweep_tile.log
  SOC_AHEAD=$a timeout 600 python tools/sweep.py --deposit 2 --refill 8 --agg 6,24 --reps 3 >> gpurun_out/sweep_tile.log 2>&1
done
cut -c1-200 gpurun_out/sweep_tile.log
