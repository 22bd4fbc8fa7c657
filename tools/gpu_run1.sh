set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lean.log
tail -5 gpurun_out/pytest_lean.log
SOC_LAYOUT=1 python bench.py --steps 3 --warmup 3 --no-cpu --kernel-times > gpurun_out/bench_lean_brick.log 2>&1; tail -1 gpurun_out/bench_lean_brick.log
SOC_LAYOUT=0 python bench.py --steps 3 --warmup 3 --no-cpu --kernel-times > gpurun_out/bench_lean_linear.log 2>&1; tail -1 gpurun_out/bench_lean_linear.log
