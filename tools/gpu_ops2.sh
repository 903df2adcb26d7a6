#!/bin/bash
# pairalign -a over two devices: same op strings / same text as one device, and the timing at config 5 sizes
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "two_devices" 2>&1 | tail -5
echo "== CTA kernel, 1184 pairs, devices 0,1"
timeout 300 python tools/ops_bench.py --pairs 1184 --devices 0,1 --tag cta2
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_long(200, 1005)
synth.write_fasta("/tmp/c5.fst", names, seqs)
names, seqs = synth.make_16s_like(300, 1002)
synth.write_fasta("/tmp/c2s.fst", names, seqs)
PY
echo "== command line, config 5, -a -n, 2 devices (one device: 9fdbc979ac14b6ca303f83d01628d930, 20.8 s)"
( time PAIRALIGN_TIMING=1 timeout 600 build/pairalign_b200 -a -n /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -14
echo "== command line, 300 x 1.5 kb, -a -n, 1 and 2 devices"
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 build/pairalign_b200 -a -n /tmp/c2s.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -6
( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0,1 build/pairalign_b200 -a -n /tmp/c2s.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -6
} 2>&1 | tee gpurun_out/ops2.log
