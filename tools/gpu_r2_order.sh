#!/bin/bash
L=gpurun_out/r2_order.log
: > $L
for o in 0 1 0 1; do
  echo "== 256^3 SOC_UNIT_ORDER=$o" >> $L
  SOC_UNIT_ORDER=$o python tools/sweep.py --n 256 --reps 3 --deposit 2 >> $L 2>&1
done
cat $L
