#!/bin/bash
set -x
mkdir -p gpurun_out
for C in 16 32; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nj_sums --launch-skip 64 --launch-count 1 -f -o gpurun_out/nj_sums_r8000_c$C python tools/nj_bench.py --taxa 8000 --ref-taxa 0 --cols $C > gpurun_out/nj_ncu.log 2>&1
ncu -i gpurun_out/nj_sums_r8000_c$C.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__inst_executed.sum,launch__grid_size > gpurun_out/nj_sums_c${C}_raw.csv 2>&1
tail -1 gpurun_out/nj_sums_c${C}_raw.csv | cut -c1-400
done
