for nbr in 1 0; do SOC_NBR=$nbr python tools/bench_octree.py --cpu-seconds 0.5 --bg-batch 100 2>&1 | grep sim_walk | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NBR=$nbr batch100', d['cell_steps_per_s'], d['packets_per_s'], d['ms'], d['stuck'])"; done
ncu --set full --clock-control none --import-source on -k regex:sim_link_kernel -s 1 -c 1 -o gpurun_out/prof_r1_link -f python tools/bench_octree.py --cpu-seconds 0.2 --bg-batch 60 > gpurun_out/ncu_link.log 2>&1
ncu -i gpurun_out/prof_r1_link.ncu-rep --page details > gpurun_out/prof_r1_link_details.txt 2>&1
