#!/bin/bash
# Command line at BASELINE.json config-2 size: wall time and phase split, 1 and 2 devices.
mkdir -p gpurun_out
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_16s_like(1000, 1002)
synth.write_fasta("/tmp/c2.fst", names, seqs)
PY
python -m phylommand_b200.build > /dev/null
for dev in 0 0,1; do
  s=$(date +%s.%N)
  PAIRALIGN_DEVICES=$dev build/pairalign_b200 -j -n -m /tmp/c2.fst > /tmp/c2_$dev.out
  e=$(date +%s.%N)
  echo "cli c2 -j -n -m devices=$dev: $(echo "$e - $s" | bc -l 2>/dev/null || python -c "print($e-$s)") s, $(md5sum < /tmp/c2_$dev.out)" | tee -a gpurun_out/cli_scale.log
done
s=$(date +%s.%N); PAIRALIGN_DEVICES=0 build/pairalign_b200 -d -n /tmp/c2.fst > /tmp/c2_d.out; e=$(date +%s.%N)
echo "cli c2 -d -n devices=0: $(python -c "print($e-$s)") s, $(wc -c < /tmp/c2_d.out) bytes" | tee -a gpurun_out/cli_scale.log
