#!/bin/bash
L=gpurun_out/r2_dom_sweep6.log
: > $L
for chunk in 16777216 33554432; do for cl in 131072 524288; do for srt in 0 1 2 3; do
  echo "== 512^3 chunk=$chunk cleanup=$cl sort=$srt" >> $L
  SOC_DOMAIN_CHUNK=$chunk SOC_DOMAIN_CLEANUP=$cl SOC_DOMAIN_SORT=$srt python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
done; done; done
echo "== verbose PS" >> $L
SOC_DOMAIN_VERBOSE=2 SOC_DOMAIN_CHUNK=33554432 SOC_DOMAIN_SORT=1 python tools/sweep.py --n 512 --reps 1 --deposit 2 2>&1 | grep "domain" | head -60 >> $L
cat $L
