#!/bin/bash
# round 2: domain-tiled propagation -- equality tests, then timings at 512^3 and 256^3 with and without domains
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "domain_tiled" > gpurun_out/r2_dom_tests.log 2>&1
tail -5 gpurun_out/r2_dom_tests.log
export SOC_DOMAIN_VERBOSE=1
for dom in -1 0; do
  echo "== 512^3 SOC_DOMAINS=$dom" >> gpurun_out/r2_dom_sweep.log
  SOC_DOMAINS=$dom timeout 600 python tools/sweep.py --n 512 --reps 2 --deposit 2 >> gpurun_out/r2_dom_sweep.log 2>&1
done
for dom in -1 256 128; do
  echo "== 256^3 SOC_DOMAINS=$dom" >> gpurun_out/r2_dom_sweep.log
  SOC_DOMAINS=$dom timeout 600 python tools/sweep.py --n 256 --reps 2 --deposit 2 >> gpurun_out/r2_dom_sweep.log 2>&1
done
echo "== 512^3 with_abu SOC_DOMAINS=0" >> gpurun_out/r2_dom_sweep.log
SOC_DOMAINS=0 timeout 600 python tools/sweep.py --n 512 --reps 2 --deposit 2 --opts noabsorbed=0,with_abu=1 >> gpurun_out/r2_dom_sweep.log 2>&1
cat gpurun_out/r2_dom_sweep.log
