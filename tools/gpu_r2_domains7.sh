#!/bin/bash
L=gpurun_out/r2_dom_sweep7.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "domain_tiled" > gpurun_out/r2_dom_tests7.log 2>&1
tail -3 gpurun_out/r2_dom_tests7.log
: > $L
for ctas in 3 4; do for srt in 0 1; do
  echo "== 512^3 ctas=$ctas sort=$srt" >> $L
  SOC_DOM_CTAS=$ctas SOC_DOMAIN_SORT=$srt python tools/sweep.py --n 512 --reps 3 --deposit 2 >> $L 2>&1
done; done
echo "== 512^3 abu ctas=4" >> $L
python tools/sweep.py --n 512 --reps 2 --deposit 2 --opts noabsorbed=0,with_abu=1 >> $L 2>&1
echo "== verbose" >> $L
SOC_DOMAIN_VERBOSE=2 python tools/sweep.py --n 512 --reps 1 --deposit 2 2>&1 | grep "domain" | head -24 >> $L
cat $L
