#!/bin/bash
# end-of-round evidence: int32 kernels with immediate gap penalties, config 4 pipeline busy times, ncu of the final s16x2 kernel
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or long or alignments or lists or four or scoring or traceback" 2>&1 | tail -4
echo "== -a, CTA kernel, 592 pairs of 30 kb"
timeout 300 python tools/ops_bench.py --pairs 592 --tag cta-imm
echo "== statistics, CTA kernel forced, 16 x 30 kb (120 pairs)"
PAIRALIGN_FORCE_CTA=1 timeout 600 python bench.py --workload c5s --steps 2 --warmup 2 --no-cpu-baseline --no-peak 2>&1 | tail -1 | cut -c1-330
python - <<'PY'
from phylommand_b200 import synth
names, seqs, taxa = synth.make_its_like(5000, 1004)
synth.write_fasta("/tmp/c4.fst", names, seqs, taxa=taxa)
PY
for mode in "both:cut-off=0.97" "alignment_groups"; do
  echo "== config 4, --group $mode"
  ( time PAIRALIGN_TIMING=1 PAIRALIGN_DEVICES=0 build/pairalign_b200 --group $mode /tmp/c4.fst | md5sum ) 2>&1 | grep -a "pipeline busy\|align + replay\|real\|  -"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v5.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_warp_duo -s 3 -c 1 -f -o gpurun_out/prof_v5_duo \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
} 2>&1 | tee gpurun_out/r1e.log
