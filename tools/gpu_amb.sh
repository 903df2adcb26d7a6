#!/bin/bash
mkdir -p gpurun_out
for w in c2n c2 c4; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-peak 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', round(d['gcups'],1), round(d['ms_per_step'],1), d['parity_spot_check'], d['gpu_launches'])"
done | tee gpurun_out/amb_bench.log
timeout 1200 python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
