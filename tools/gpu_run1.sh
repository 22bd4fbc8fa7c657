python -m pytest tests/test_gpu_parity.py tests/test_driver_gpu.py -m gpu -q -k "statistical or driver or sharded or invariants" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|assert|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -20
for nbr in 1 0; do SOC_NBR=$nbr python tools/bench_octree.py --cpu-seconds 0.5 2>&1 | grep sim_walk | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NBR=$nbr', d['cell_steps_per_s'], d['packets_per_s'], d['ms'], d['stuck'])"; done
