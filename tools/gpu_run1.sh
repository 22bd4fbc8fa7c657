python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roi" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|assert|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -30
