#!/bin/bash
# First-contact GPU session: parity tests, smoke, instruction peaks, a short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
cat gpurun_out/smoke.log
timeout 300 python - > gpurun_out/peaks.log 2>&1 <<'PY'
from phylommand_b200 import capi
capi.init([0])
for w, name in capi.PEAK_CLASSES.items():
    g, mhz = capi.int32_peak(w)
    print(f"{w:2d} {name:22s} {g:10.1f} Gop/s  clk {mhz:7.1f} MHz  -> {g*1e3/mhz/148:6.1f} lanes/clk/SM")
capi.shutdown()
PY
cat gpurun_out/peaks.log
timeout 900 python bench.py --workload "${BENCH_WORKLOAD:-c2}" --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1
cat gpurun_out/bench.log
