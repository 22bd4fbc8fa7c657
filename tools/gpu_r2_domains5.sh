#!/bin/bash
L=gpurun_out/r2_dom_sweep5.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "domain_tiled" > gpurun_out/r2_dom_tests5.log 2>&1
tail -5 gpurun_out/r2_dom_tests5.log
: > $L
for srt in 0 1; do
  echo "== 512^3 SOC_DOMAIN_SORT=$srt" >> $L
  SOC_DOMAIN_VERBOSE=1 SOC_DOMAIN_SORT=$srt python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
done
echo "== 512^3 sort, cleanup 4M" >> $L
SOC_DOMAIN_CLEANUP=4000000 python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
echo "== 512^3 sort, cleanup 256k" >> $L
SOC_DOMAIN_CLEANUP=256000 python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
echo "== 256^3 SOC_DOMAINS=256 sorted" >> $L
SOC_DOMAINS=256 python tools/sweep.py --n 256 --reps 2 --deposit 2 >> $L 2>&1
echo "== 256^3 SOC_DOMAINS=128 sorted" >> $L
SOC_DOMAINS=128 python tools/sweep.py --n 256 --reps 2 --deposit 2 >> $L 2>&1
grep -v "soc_b200:   domain" $L
