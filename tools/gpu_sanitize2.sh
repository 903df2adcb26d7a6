#!/bin/bash
# compute-sanitizer over what changed late in round 1: the biased s16x2 kernel at its limits, the CTA move-storing
# kernel, the warp-per-pair walk with its shared-memory windows (memcheck; racecheck for the walk's double buffer)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "batched or traceback or near_its_int16 or strip_width or floating_window or sparse_ambiguity" > gpurun_out/sanitizer2.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "test_batched_alignments and not long" >> gpurun_out/sanitizer2.log 2>&1
echo "racecheck walk rc=$?" | tee -a gpurun_out/sanitizer2.log
grep -a "passed\|failed\|ERROR SUMMARY\|rc=\|RACECHECK SUMMARY" gpurun_out/sanitizer2.log
