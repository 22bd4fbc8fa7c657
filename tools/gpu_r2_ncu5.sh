#!/bin/bash
O=gpurun_out
export SOC_TWO_PASS=1 SOC_TILEPASS_REFILL=24 SOC_TILEPASS_AGG=2
python tools/prof_two_pass.py > $O/r2m_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sim_tile_pass_kernel -s 2 -c 1 -f -o $O/r2m_prof_tile python tools/prof_two_pass.py > $O/r2m_ncu.log 2>&1
ncu -i $O/r2m_prof_tile.ncu-rep --page details > $O/r2m_prof_tile_details.txt 2>&1
ncu -i $O/r2m_prof_tile.ncu-rep --page source --csv > $O/r2m_prof_tile_source.csv 2>&1
rm -f $O/r2m_prof_tile.ncu-rep
grep -E "^\s+(Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|L2 Hit|DRAM Throughput|L2 Cache Throughput|Avg. Active Threads|No Eligible|Grid Size)" $O/r2m_prof_tile_details.txt
