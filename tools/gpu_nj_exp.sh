#!/bin/bash
mkdir -p gpurun_out
for st in 3 6 12 24; do for pr in 0 1; do
echo "stages=$st promo=$pr"; PAIRALIGN_NJ_STAGES=$st PAIRALIGN_NJ_PROMO=$pr timeout 300 python tools/nj_bench.py --taxa 8000 --ref-taxa 0 2>&1 | tail -1
done; done | tee gpurun_out/nj_exp.log
