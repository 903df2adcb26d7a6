#!/bin/bash
# pairalign -a on long pairs: parity of the CTA move-storing kernel, then timings at config 5 sizes
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alignments or traceback or cta" 2>&1 | tail -15
echo "== CTA kernel, 592 pairs"
timeout 300 python tools/ops_bench.py --pairs 592 --tag cta
echo "== one pair per warp (PAIRALIGN_NO_CTA=1), 296 pairs"
PAIRALIGN_NO_CTA=1 timeout 300 python tools/ops_bench.py --pairs 296 --tag warp
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_long(200, 1005)
synth.write_fasta("/tmp/c5.fst", names, seqs)
PY
echo "== command line, config 5, -a -n"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 timeout 600 build/pairalign_b200 -a -n /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -14
echo "== command line, config 5, -p -m"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 timeout 600 build/pairalign_b200 -p -m /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -14
} 2>&1 | tee gpurun_out/ops.log
