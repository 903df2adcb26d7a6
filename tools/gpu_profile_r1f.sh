#!/bin/bash
# ncu captures of the final walk kernel (shared-memory windows) and the general IUPAC / gap kernel
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_walk -s 1 -c 1 -f -o gpurun_out/prof_walk2 \
    python tools/ops_bench.py --seqs 40 --pairs 148 > gpurun_out/ncu_walk2.log 2>&1
PAIRALIGN_NO_AMB=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_warp_dp_kernel -s 1 -c 1 -f -o gpurun_out/prof_general \
    python bench.py --workload c2n --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/ncu_general.log 2>&1
tail -n 2 gpurun_out/ncu_general.log | cut -c1-200
