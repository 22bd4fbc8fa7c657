#!/bin/bash
L=gpurun_out/r2_dom_sweep3.log
: > $L
for cfg in "256 256 512" "512 256 256" "256 512 256"; do
  for dom in -1 256; do
    echo "== $cfg SOC_DOMAINS=$dom" >> $L
    SOC_DOMAIN_VERBOSE=1 SOC_DOMAINS=$dom python tools/box_sweep.py $cfg 8 >> $L 2>&1
  done
done
for dom in -1 0; do
  echo "== 512^3 SOC_DOMAINS=$dom" >> $L
  SOC_DOMAIN_VERBOSE=1 SOC_DOMAINS=$dom python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
done
echo "== 256^3 SOC_DOMAINS=256" >> $L
SOC_DOMAINS=256 python tools/sweep.py --n 256 --reps 2 --deposit 2 >> $L 2>&1
grep -v "soc_b200: " $L
