#!/bin/bash
# NJ session: GPU tests (NJ only), NJ timings, ncu over the first rounds of an 8000-taxa tree.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nj.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_nj.log
cat gpurun_out/pytest_gpu_nj.log
for c in 0; do timeout 300 python tools/nj_bench.py --taxa 1000,3000,6000,10000 --ref-taxa 0 --cols $c > gpurun_out/nj_bench_cols$c.log 2>&1; cat gpurun_out/nj_bench_cols$c.log; done
NJ_COLS=0 bash tools/gpu_nj_prof.sh
