#!/bin/bash
# ncu captures of the long-pair kernels on a 16 x 30 kb set (120 pairs).
set -x
mkdir -p gpurun_out
PAIRALIGN_FORCE_CTA=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_cta32 -s 1 -c 1 -f -o gpurun_out/prof_cta32 \
    python bench.py --workload c5s --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/ncu_cta.log 2>&1
PAIRALIGN_NO_CTA=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_warp32 -s 1 -c 1 -f -o gpurun_out/prof_warp32 \
    python bench.py --workload c5s --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/ncu_warp32.log 2>&1
for v in FORCE_CTA NO_CTA; do env PAIRALIGN_$v=1 timeout 600 python bench.py --workload c5s --steps 2 --warmup 2 --no-cpu-baseline --no-peak 2>&1 | tail -1 | cut -c1-330; done
