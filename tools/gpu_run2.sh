set -x
# launch list (all kernels of the bench step) and one full capture of the background launch
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_lean.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_lean_kernel -s 7 -c 1 -o gpurun_out/prof_r1_lean_bg -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/prof_r1_lean_bg.ncu-rep --page details > gpurun_out/prof_r1_lean_bg_details.txt 2>&1
tail -3 gpurun_out/ncu_full2.log
