#!/bin/bash
L=gpurun_out/r2_dom_sweep2.log
echo "== 256^3 scrambled order" > $L
SOC_SCRAMBLE=1 python tools/sweep.py --n 256 --reps 2 --deposit 2 >> $L 2>&1
echo "== 512^3 scrambled order" >> $L
SOC_SCRAMBLE=1 SOC_DOMAINS=-1 python tools/sweep.py --n 512 --reps 2 --deposit 2 >> $L 2>&1
echo "== 512^3 domains verbose 2" >> $L
SOC_DOMAIN_VERBOSE=2 SOC_DOMAINS=0 python tools/sweep.py --n 512 --reps 1 --deposit 2 >> $L 2>&1
cat $L | tail -150
