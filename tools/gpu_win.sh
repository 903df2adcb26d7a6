#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_win.log
cat gpurun_out/pytest_gpu_win.log
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/bench_c5_win.log 2>&1
tail -1 gpurun_out/bench_c5_win.log | cut -c1-1400
