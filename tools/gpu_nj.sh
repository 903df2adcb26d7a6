#!/bin/bash
# NJ session: GPU tests, timings (with the reference treeator on 1500 taxa), ncu of one nj_sums / nj_argmin launch.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nj.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_nj.log
cat gpurun_out/pytest_gpu_nj.log
timeout 600 python tools/nj_bench.py --taxa 1000,3000,6000,10000 --ref-taxa 1500 > gpurun_out/nj_bench.log 2>&1
cat gpurun_out/nj_bench.log
bash tools/gpu_nj_prof.sh
