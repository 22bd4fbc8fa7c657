#!/bin/bash
# ncu evidence for the domain kernels at 512^3 after the staged hand-over: metric list over every launch of the second
# background launch, and one full capture of the first (largest) domain launch.
O=gpurun_out
SOC_DOMAIN_VERBOSE=2 python tools/prof_512.py > $O/r2e_plain_512.log 2>&1 || exit 1
ncu --clock-control none --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio \
    -k regex:"sim_|q_|fold" -c 400 --csv --log-file $O/r2e_launches_512.csv python tools/prof_512.py > $O/r2e_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 42 -c 1 -f -o $O/r2e_prof_ahead_512 python tools/prof_512.py > $O/r2e_ncu_512.log 2>&1
ncu -i $O/r2e_prof_ahead_512.ncu-rep --page details > $O/r2e_prof_ahead_512_details.txt 2>&1
ncu -i $O/r2e_prof_ahead_512.ncu-rep --page raw --csv > $O/r2e_prof_ahead_512_raw.csv 2>&1
ncu -i $O/r2e_prof_ahead_512.ncu-rep --page source --csv > $O/r2e_prof_ahead_512_source.csv 2>&1
rm -f $O/r2e_prof_ahead_512.ncu-rep
grep -E "^\s+(Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|L2 Hit|DRAM Throughput|L2 Cache Throughput|Avg. Active Threads|No Eligible|Grid Size)" $O/r2e_prof_ahead_512_details.txt
