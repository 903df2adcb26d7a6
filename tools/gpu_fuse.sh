#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for w in c2 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-peak 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['gcups'], d['ms_per_step'], d['parity_spot_check'])"
done
