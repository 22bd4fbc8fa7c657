ncu --set full --clock-control none --import-source on -k regex:sim_link_kernel -s 1 -c 1 -o gpurun_out/prof_r1_link -f python tools/bench_octree.py --cpu-seconds 0.2 --bg-batch 60 > gpurun_out/ncu_link.log 2>&1
ncu -i gpurun_out/prof_r1_link.ncu-rep --page details > gpurun_out/prof_r1_link_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sca_link_kernel -s 1 -c 1 -o gpurun_out/prof_r1_scalink -f python tools/bench_octree.py --cpu-seconds 0.2 --sca-packets 2e6 > gpurun_out/ncu_scalink.log 2>&1
ncu -i gpurun_out/prof_r1_scalink.ncu-rep --page details > gpurun_out/prof_r1_scalink_details.txt 2>&1
grep -h "link_kernel\|Duration\|L2 Cache Throughput\|Issue Slots Busy\|L1/TEX Hit\|L2 Hit\|Registers Per\|Avg. Active Threads\|Not Predicated" gpurun_out/prof_r1_link_details.txt gpurun_out/prof_r1_scalink_details.txt
