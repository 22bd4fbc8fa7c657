#!/bin/bash
L=gpurun_out/r2_dom_sweep4.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "domain_tiled or brick_layout or kernel_variants or accumulation_engines" > gpurun_out/r2_dom_tests4.log 2>&1
tail -5 gpurun_out/r2_dom_tests4.log
: > $L
for dom in -1 0; do
  echo "== 512^3 SOC_DOMAINS=$dom" >> $L
  SOC_DOMAIN_VERBOSE=1 SOC_DOMAINS=$dom python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
done
echo "== 512^3 SOC_DOMAINS=0 verbose 2 (one rep)" >> $L
SOC_DOMAIN_VERBOSE=2 SOC_DOMAINS=0 python tools/sweep.py --n 512 --reps 1 --deposit 2 2>&1 | tail -120 >> $L
echo "== 512^3 with_abu SOC_DOMAINS=0" >> $L
SOC_DOMAINS=0 python tools/sweep.py --n 512 --reps 2 --deposit 2 --opts noabsorbed=0,with_abu=1 >> $L 2>&1
echo "== 256^3 default" >> $L
python tools/sweep.py --n 256 --reps 2 --deposit 2 >> $L 2>&1
grep -v "soc_b200:   domain" $L
