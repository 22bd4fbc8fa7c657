SOC_AHEAD=3 ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 6 -c 1 -o gpurun_out/prof_r1_ahead_ps -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_aps.log 2>&1
ncu -i gpurun_out/prof_r1_ahead_ps.ncu-rep --page details > gpurun_out/prof_r1_ahead_ps_details.txt 2>&1
grep -h "sim_ahead\|Duration\|L2 Cache Throughput\|Issue Slots Busy\|L1/TEX Hit\|L2 Hit\|Registers Per\|Avg. Active Threads" gpurun_out/prof_r1_ahead_ps_details.txt
