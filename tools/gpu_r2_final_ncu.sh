#!/bin/bash
# final round-2 ncu session: launch list of the bench step and full captures of its three packet kernels (tile pass, queue-fed
# second pass of the point-source launch, background launch).  Reports are turned into text here and removed.
O=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-extras --no-driver"
export_rep() {   # $1 = report base name
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>&1
  ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2>&1
  rm -f $O/$1.ncu-rep
}
$BENCH > $O/r2f_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2f_launches_bench.csv $BENCH > $O/r2f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sim_tile_pass_kernel -s 1 -c 1 -f -o $O/r2f_prof_tile_pass $BENCH > $O/r2f_ncu_tile.log 2>&1
export_rep r2f_prof_tile_pass
# the look-ahead kernel runs twice per step: queue-fed second pass of the point-source launch, then the background launch
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 2 -c 1 -f -o $O/r2f_prof_ahead_a $BENCH > $O/r2f_ncu_a.log 2>&1
export_rep r2f_prof_ahead_a
ncu --set full --clock-control none --import-source on -k regex:sim_ahead_kernel -s 3 -c 1 -f -o $O/r2f_prof_ahead_b $BENCH > $O/r2f_ncu_b.log 2>&1
export_rep r2f_prof_ahead_b
for f in tile_pass ahead_a ahead_b; do
  echo "== $f: $(grep -m1 -E 'sim_[a-z_]+kernel' $O/r2f_prof_${f}_details.txt | cut -c1-120)"
  grep -E "^\s+(Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|L2 Hit|DRAM Throughput|L2 Cache Throughput|Avg. Active Threads|No Eligible|Grid Size)" $O/r2f_prof_${f}_details.txt
  python - <<PY
import csv
rows=list(csv.reader(open("$O/r2f_prof_${f}_raw.csv")))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
for k in ("dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum","smsp__inst_executed.sum"):
    if k in h: print("   ",k, v[h.index(k)], rows[1][h.index(k)])
PY
done
