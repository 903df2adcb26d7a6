#!/bin/bash
# what the driver runs at round end: the GPU tests, smoke(), both bench arms
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-600
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-700 gpurun_out/bench_default.json
} 2>&1 | tee gpurun_out/final.log
