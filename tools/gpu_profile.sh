#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the DP kernel.
set -x
mkdir -p gpurun_out
W="${BENCH_WORKLOAD:-tiny}"
TAG="${PROFILE_TAG:-dp}"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --workload "$W" --steps 2 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_warp_ -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --workload "$W" --steps 1 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
