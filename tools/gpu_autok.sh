#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1 set=$2 a=$3"; PAIRALIGN_KDUO_SET=$2 PAIRALIGN_KDUO_A=$3 timeout 600 python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu-baseline --no-peak 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['gcups'], d['ms_per_step'])"; }
{
run c4 0x3e00 256     # 9..13
run c4 0x3d00 256     # 8,10..13
run c4 0x3d80 256     # 7,8,10..13
run c4 0x3f00 256     # 8..13
run c4 0x3f00 100
} 2>&1 | tee gpurun_out/autok_sweep2.log
