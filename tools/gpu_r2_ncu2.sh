#!/bin/bash
O=gpurun_out
python tools/prof_512.py > $O/r2_plain_512b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 42 -c 1 -f -o $O/r2_prof_ahead_512_first python tools/prof_512.py > $O/r2_ncu_512b.log 2>&1
ncu -i $O/r2_prof_ahead_512_first.ncu-rep --page details > $O/r2_prof_ahead_512_first_details.txt 2>&1
ncu -i $O/r2_prof_ahead_512_first.ncu-rep --page raw --csv > $O/r2_prof_ahead_512_first_raw.csv 2>&1
ncu -i $O/r2_prof_ahead_512_first.ncu-rep --page source --csv > $O/r2_prof_ahead_512_first_source.csv 2>&1
rm -f $O/r2_prof_ahead_512_first.ncu-rep
grep -E "^\s+(Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|L2 Hit|DRAM Throughput|L2 Cache Throughput|Avg. Active Threads|No Eligible|Grid Size)" $O/r2_prof_ahead_512_first_details.txt
