#!/bin/bash
# compute-sanitizer memcheck over the small parity tests (every kernel family incl. the strip-width variants, the
# floating-window s16x2 path and neighbour joining) and racecheck over the NJ kernels (shared-memory rings).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "golden or mixed or sub_ranges or aligned_mode or other_scoring or batched or traceback or strip_width or floating_window" > gpurun_out/sanitizer.log 2>&1
echo "memcheck parity rc=$?" | tee -a gpurun_out/sanitizer.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_nj.py -m gpu -x -q \
   -k "matches_oracle and not 1500 and not 2000 and not 1000" >> gpurun_out/sanitizer.log 2>&1
echo "memcheck nj rc=$?" | tee -a gpurun_out/sanitizer.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_nj.py -m gpu -x -q \
   -k "matches_oracle and (ties-257 or rand-33 or jc-200)" >> gpurun_out/sanitizer.log 2>&1
echo "racecheck nj rc=$?" | tee -a gpurun_out/sanitizer.log
grep -a "passed\|failed\|ERROR SUMMARY\|rc=\|RACECHECK SUMMARY" gpurun_out/sanitizer.log
