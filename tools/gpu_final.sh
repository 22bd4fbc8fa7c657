set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
python tools/bench_octree.py > gpurun_out/octree.log 2>&1; tail -3 gpurun_out/octree.log | cut -c1-300
python tools/sweep.py --n 512 --deposit 2 --refill 8 --agg 24 --reps 2 > gpurun_out/sweep_512.log 2>&1; tail -1 gpurun_out/sweep_512.log
python tools/sweep.py --n 512 --deposit 2 --refill 8 --agg 24 --reps 2 --opts noabsorbed=0,with_abu=1 > gpurun_out/sweep_512_abu.log 2>&1; tail -1 gpurun_out/sweep_512_abu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 3 -c 1 -o gpurun_out/prof_r1_ahead_bg -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_bg.log 2>&1
ncu -i gpurun_out/prof_r1_ahead_bg.ncu-rep --page details > gpurun_out/prof_r1_ahead_bg_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_lean_kernel -s 3 -c 1 -o gpurun_out/prof_r1_lean_ps -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_ps.log 2>&1
ncu -i gpurun_out/prof_r1_lean_ps.ncu-rep --page details > gpurun_out/prof_r1_lean_ps_details.txt 2>&1
rm -f gpurun_out/prof_r1_lean_ps.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:sim_link_kernel -s 1 -c 1 -o gpurun_out/prof_r1_link -f python tools/bench_octree.py --cpu-seconds 0.2 --bg-batch 60 > gpurun_out/ncu_link.log 2>&1
ncu -i gpurun_out/prof_r1_link.ncu-rep --page details > gpurun_out/prof_r1_link_details.txt 2>&1
rm -f gpurun_out/prof_r1_link.ncu-rep
