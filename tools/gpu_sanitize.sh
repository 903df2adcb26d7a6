#!/bin/bash
# compute-sanitizer memcheck over the small parity tests (every kernel family), then config 3 once.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "golden or mixed or sub_ranges or aligned_mode or other_scoring or batched or traceback" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" | tee -a gpurun_out/sanitizer.log
tail -12 gpurun_out/sanitizer.log
timeout 900 python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu-baseline --no-peak > gpurun_out/bench_c3.log 2>&1
tail -1 gpurun_out/bench_c3.log | cut -c1-1200
