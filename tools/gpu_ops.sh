#!/bin/bash
# pairalign -a: parity of the move-storing kernels and the walk, then timings at config 5 sizes
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alignments or traceback or cta" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "_a" 2>&1 | tail -3
echo "== CTA kernel, 592 pairs"
timeout 300 python tools/ops_bench.py --pairs 592 --tag cta
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_long(200, 1005)
synth.write_fasta("/tmp/c5.fst", names, seqs)
names, seqs = synth.make_16s_like(300, 1002)
synth.write_fasta("/tmp/c2s.fst", names, seqs)
PY
echo "== command line, config 5, -a -n   (expected md5 9fdbc979ac14b6ca303f83d01628d930)"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 timeout 600 build/pairalign_b200 -a -n /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -14
echo "== command line, 300 x 1.5 kb, -a -n   (expected md5 52f6df747d60591d05b8d42a630f5376)"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 build/pairalign_b200 -a -n /tmp/c2s.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -6
} 2>&1 | tee gpurun_out/ops.log
