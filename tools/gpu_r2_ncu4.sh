#!/bin/bash
O=gpurun_out
export SOC_TWO_PASS=1
for v in 0 1; do
  export SOC_TILE_PASS=$v
  python tools/prof_two_pass.py > $O/r2j_plain_$v.log 2>&1 || exit 1
  K=sim_lean_kernel; [ $v = 1 ] && K=sim_tile_pass_kernel
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $O/r2j_prof_tile_$v python tools/prof_two_pass.py > $O/r2j_ncu_$v.log 2>&1
  ncu -i $O/r2j_prof_tile_$v.ncu-rep --page details > $O/r2j_prof_tile_${v}_details.txt 2>&1
  ncu -i $O/r2j_prof_tile_$v.ncu-rep --page source --csv > $O/r2j_prof_tile_${v}_source.csv 2>&1
  rm -f $O/r2j_prof_tile_$v.ncu-rep
  grep -E "^\s+(Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|L2 Hit|DRAM Throughput|L2 Cache Throughput|Avg. Active Threads|No Eligible|Grid Size)" $O/r2j_prof_tile_${v}_details.txt
done
