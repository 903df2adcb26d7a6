#!/bin/bash
# general (IUPAC / gap) kernel: parity, then its rate on config 2 with two ambiguity codes per sequence and the
# s16x2 AMB variant switched off (every pair takes the general kernel)
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "example or mixed or aligned" 2>&1 | tail -3
echo "== c2n, PAIRALIGN_NO_AMB=1 (before: 690 GCUPS)"
PAIRALIGN_NO_AMB=1 timeout 600 python bench.py --workload c2n --steps 2 --warmup 1 --no-cpu-baseline --no-peak 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d[k] for k in ('value','ms_per_step','gcups','parity_spot_check','gpu_launches')})"
} 2>&1 | tee gpurun_out/gen.log
