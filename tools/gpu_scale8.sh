#!/bin/bash
# 8 x B200: config 3 (10 000 x 1.5 kb) with the set fixed (strong scaling), and the command line on 8 devices
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload c3 --scaling strong --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/bench_c3_8gpu_strong.log 2>&1
tail -1 gpurun_out/bench_c3_8gpu_strong.log | cut -c1-1200
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_16s_like(1000, 1002)
synth.write_fasta("/tmp/c2.fst", names, seqs)
PY
for dev in 0 0,1,2,3,4,5,6,7; do
  ( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=$dev build/pairalign_b200 -j -n -m /tmp/c2.fst > /tmp/c2_out.txt ) 2>&1 | tail -8
  md5sum /tmp/c2_out.txt
done 2>&1 | tee gpurun_out/cli_8gpu.log
