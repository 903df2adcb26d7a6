#!/bin/bash
# Other workloads and the command line at scale (2 GPUs visible).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/bench_c4.log 2>&1
tail -1 gpurun_out/bench_c4.log | cut -c1-700
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-peak > gpurun_out/bench_c5.log 2>&1
tail -1 gpurun_out/bench_c5.log | cut -c1-700
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_16s_like(1000, 1002)
synth.write_fasta("/tmp/c2.fst", names, seqs)
names, seqs, taxa = synth.make_its_like(1500, 1004)
synth.write_fasta("/tmp/c4s.fst", names, seqs, taxa=taxa)
PY
python -m phylommand_b200.build
for dev in 0 0,1; do
  /usr/bin/env time -f "cli c2 -j -n -m devices=$dev: %e s wall" env PAIRALIGN_DEVICES=$dev build/pairalign_b200 -j -n -m /tmp/c2.fst > /tmp/c2_$dev.out 2> gpurun_out/cli_c2_$dev.err || true
  tail -1 gpurun_out/cli_c2_$dev.err
  md5sum /tmp/c2_$dev.out
done
( time PAIRALIGN_DEVICES=0 build/pairalign_b200 -g both:cut-off=0.97 /tmp/c4s.fst > gpurun_out/cli_c4s_groups.out ) 2>&1 | tail -4
head -c 300 gpurun_out/cli_c4s_groups.out
# the first 64 sequences against the reference binary, byte for byte
python - <<'PY'
from phylommand_b200 import synth
names, seqs = synth.make_16s_like(1000, 1002)
synth.write_fasta("/tmp/c2_64.fst", names[:64], seqs[:64])
PY
( time oracle/_ref/pairalign_pthread -T 16 -j -n -m /tmp/c2_64.fst > /tmp/ref64.out ) 2>&1 | tail -3
oracle/_ref/pairalign -j -n -m /tmp/c2_64.fst > /tmp/ref64_st.out
build/pairalign_b200 -j -n -m /tmp/c2_64.fst > /tmp/ours64.out
cmp /tmp/ref64_st.out /tmp/ours64.out && echo "C2 prefix (64 seqs, 2016 pairs): stdout byte-identical to the reference"
