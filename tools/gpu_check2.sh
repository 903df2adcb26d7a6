#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r01_c2.json 2> gpurun_out/bench_r01_c2.err
cat gpurun_out/bench_r01_c2.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; cut -c1-400 gpurun_out/bench_ref.log
timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/bench_c5.log 2>&1
tail -1 gpurun_out/bench_c5.log | cut -c1-330
