#!/bin/bash
# the default bench line at N GPUs, as the driver launches it (weak scaling)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_${N}gpu_weak.log 2>&1
tail -1 gpurun_out/bench_c2_${N}gpu_weak.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'GPUs:', round(d['gcups'],1), 'GCUPS', round(d['value']), 'pairs/s', round(d['ms_per_step'],1), 'ms/step; e2e', round(d['e2e']['value']), 'pairs/s; roofline frac', round(d['roofline']['frac'],3))"
