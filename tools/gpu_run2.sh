set -x
python bench.py > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
python tools/bench_octree.py > gpurun_out/octree.log 2>&1; tail -3 gpurun_out/octree.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_lean.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_lean_kernel -s 7 -c 1 -o gpurun_out/prof_r1_lean_bg -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/prof_r1_lean_bg.ncu-rep --page details > gpurun_out/prof_r1_lean_bg_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_link_kernel -s 1 -c 1 -o gpurun_out/prof_r1_link -f python tools/bench_octree.py --cpu-seconds 0.2 --bg-batch 60 > gpurun_out/ncu_link.log 2>&1
ncu -i gpurun_out/prof_r1_link.ncu-rep --page details > gpurun_out/prof_r1_link_details.txt 2>&1
