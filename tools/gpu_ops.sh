#!/bin/bash
# pairalign -a: parity of the move-storing kernels, timing at config 5 and config 2 sizes, command line md5
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extended.py tests/test_gpu_cli.py -m gpu -x -q -k "alignment or traceback or golden or batched or cli_matches or 30kb" 2>&1 | tail -8
echo "== CTA per item, 1184 pairs of 30 kb"
timeout 300 python tools/ops_bench.py --pairs 1184 --tag cta
echo "== warp per item, 40000 pairs of 1.5 kb"
timeout 300 python tools/ops_bench.py --seqs 300 --pairs 40000 --length 1500 --tag warp
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_long(200, 1005)
synth.write_fasta("/tmp/c5.fst", names, seqs)
PY
echo "== command line, config 5, -a -n (round 1, int32 kernels: 9fdbc979ac14b6ca303f83d01628d930, 21.2 s)"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 timeout 600 build/pairalign_b200 -a -n /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -14
} 2>&1 | tee gpurun_out/${TAG:-r02}_ops.log
