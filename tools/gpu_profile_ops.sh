#!/bin/bash
# ncu captures of the pairalign -a kernels on 148 pairs of 30 kb (one wave of the CTA kernel)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_cta32 -s 1 -c 1 -f -o gpurun_out/prof_cta32_dirs \
    python tools/ops_bench.py --seqs 40 --pairs 148 > gpurun_out/ncu_cta_dirs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_walk -s 1 -c 1 -f -o gpurun_out/prof_walk \
    python tools/ops_bench.py --seqs 40 --pairs 148 > gpurun_out/ncu_walk.log 2>&1
tail -2 gpurun_out/ncu_cta_dirs.log gpurun_out/ncu_walk.log
timeout 300 python tools/ops_bench.py --seqs 40 --pairs 148 --tag one-wave
