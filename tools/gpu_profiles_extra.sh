# extra ncu evidence: per-cell-opacity kernel (config 4 physics), scattered light on the octree (config 3), 512^3 (config 5)
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 3 -c 1 -o gpurun_out/p_kappa -f python tools/sweep.py --deposit 2 --reps 1 --opts noabsorbed=0,with_abu=1 > gpurun_out/ncu_kappa.log 2>&1
ncu -i gpurun_out/p_kappa.ncu-rep --page details > gpurun_out/prof_r1_kappa_bg_details.txt 2>&1; rm -f gpurun_out/p_kappa.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:sca_link_kernel -s 1 -c 1 -o gpurun_out/p_sca -f python tools/bench_octree.py --cpu-seconds 0.2 --sca-packets 2e6 > gpurun_out/ncu_scalink.log 2>&1
ncu -i gpurun_out/p_sca.ncu-rep --page details > gpurun_out/prof_r1_scalink_details.txt 2>&1; rm -f gpurun_out/p_sca.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 3 -c 1 -o gpurun_out/p_512 -f python tools/sweep.py --n 512 --deposit 2 --reps 1 > gpurun_out/ncu_512.log 2>&1
ncu -i gpurun_out/p_512.ncu-rep --page details > gpurun_out/prof_r1_512_bg_details.txt 2>&1; rm -f gpurun_out/p_512.ncu-rep
grep -h "ahead_kernel\|link_kernel\|Duration\|L2 Cache Throughput\|Issue Slots Busy\|DRAM Throughput\|L2 Hit" gpurun_out/prof_r1_kappa_bg_details.txt gpurun_out/prof_r1_scalink_details.txt gpurun_out/prof_r1_512_bg_details.txt
