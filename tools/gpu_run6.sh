for g in 32 64 128; do
  echo "SOC_L2_FETCH=$g" >> gpurun_out/sweep_l2fetch.log
  SOC_L2_FETCH=$g python tools/sweep.py --n 512 --deposit 2 --refill 8 --agg 24 --reps 2 >> gpurun_out/sweep_l2fetch.log 2>&1
  SOC_L2_FETCH=$g python tools/sweep.py --n 256 --deposit 2 --refill 8 --agg 24 --reps 2 >> gpurun_out/sweep_l2fetch.log 2>&1
done
cut -c1-200 gpurun_out/sweep_l2fetch.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:sim_ahead_kernel -s 1 -c 1 --csv --log-file gpurun_out/m512_bg.csv python tools/sweep.py --n 512 --deposit 2 --reps 1 > /dev/null 2>&1
grep sim_ahead gpurun_out/m512_bg.csv | cut -d, -f13- 
