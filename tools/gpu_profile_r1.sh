#!/bin/bash
# ncu evidence for profiles/ (final kernels of round 1): launch list of the default bench command, full captures of
# the s16x2 kernel on config 2 and of its floating-window variant on 64 x 7.6 kb.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_warp_duo -s 1 -c 1 -f -o gpurun_out/prof_duo_auto \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_warp_duo -s 1 -c 1 -f -o gpurun_out/prof_duo_win \
    python bench.py --workload c5w --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/ncu_win.log 2>&1
tail -2 gpurun_out/ncu_win.log | cut -c1-300
timeout 300 python bench.py --workload c5w --steps 2 --warmup 2 --no-cpu-baseline --no-peak 2>&1 | tail -1 | cut -c1-330
