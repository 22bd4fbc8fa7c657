#!/bin/bash
# round 2 ncu session: launch list of the bench step, full captures of its two kernels, DRAM traffic of a 512^3 launch.
# The .ncu-rep files are turned into text here (details / raw / source pages) and removed: gpurun_out is limited to 64 MiB.
O=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-extras"
export_rep() {   # $1 = report base name
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>&1
  ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2>&1
  rm -f $O/$1.ncu-rep
}
$BENCH > $O/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_bench.csv $BENCH > $O/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_lean_kernel -s 3 -c 1 -f -o $O/r2_prof_lean_ps $BENCH > $O/r2_ncu_ps.log 2>&1
export_rep r2_prof_lean_ps
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 3 -c 1 -f -o $O/r2_prof_ahead_bg $BENCH > $O/r2_ncu_bg.log 2>&1
export_rep r2_prof_ahead_bg
python tools/prof_512.py > $O/r2_plain_512.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file $O/r2_launches_512.csv python tools/prof_512.py > $O/r2_ncu_512_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 10 -c 1 -f -o $O/r2_prof_ahead_512_dom python tools/prof_512.py > $O/r2_ncu_512.log 2>&1
export_rep r2_prof_ahead_512_dom
SOC_DOMAINS=-1 python tools/prof_512.py > $O/r2_plain_512_whole.log 2>&1
cat $O/r2_plain_512.log $O/r2_plain_512_whole.log
du -sh $O; ls -la $O | grep r2_
