python -m pytest tests/test_gpu_parity.py tests/test_driver_gpu.py -m gpu -q -k "scattered or sca or driver" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|assert|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -20
for nbr in 1 0; do SOC_NBR=$nbr python tools/bench_octree.py --cpu-seconds 0.2 2>&1 | grep "scattered" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('NBR=$nbr', d['workload'][:22], '%.3e steps/s'%d['cell_steps_per_s'], '%.3e pk/s'%d['packets_per_s'], '%.3e peel/s'%d['peel_rays_per_s'], d['ms'], d['stuck'])"; done
