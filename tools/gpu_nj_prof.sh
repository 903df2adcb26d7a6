#!/bin/bash
# ncu over one nj_sums and one nj_argmin launch at r ~ 8000 (after the 64-taxa warm-up tree: 63 + 62 launches)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nj_sums --launch-skip 64 --launch-count 1 -f -o gpurun_out/nj_sums_r8000 python tools/nj_bench.py --taxa 8000 --ref-taxa 0 > gpurun_out/nj_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nj_argmin --launch-skip 63 --launch-count 1 -f -o gpurun_out/nj_argmin_r8000 python tools/nj_bench.py --taxa 8000 --ref-taxa 0 >> gpurun_out/nj_ncu.log 2>&1
for f in nj_sums_r8000 nj_argmin_r8000; do
ncu -i gpurun_out/$f.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__inst_executed.sum,launch__grid_size > gpurun_out/${f}_raw.csv 2>&1
cat gpurun_out/${f}_raw.csv | cut -c1-900
done
