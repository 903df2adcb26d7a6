#!/bin/bash
# config 4 through the command line: clustering and alignment groups on 5 000 ITS-like sequences
mkdir -p gpurun_out
python - <<'PY'
from phylommand_b200 import synth
names, seqs, taxa = synth.make_its_like(5000, 1004)
synth.write_fasta("/tmp/c4.fst", names, seqs, taxa=taxa)
PY
{
for mode in "both:cut-off=0.97" "alignment_groups"; do
  echo "== --group $mode"
  ( time PAIRALIGN_TIMING=1 build/pairalign_b200 --group $mode /tmp/c4.fst > /tmp/c4_$mode.out ) 2>&1 | grep -a -v "^$" | tail -12
  wc -l /tmp/c4_$mode.out; head -c 400 /tmp/c4_$mode.out; echo
done
ls -la /tmp/c4.fst*
} 2>&1 | tee gpurun_out/cli_c4_groups.log
