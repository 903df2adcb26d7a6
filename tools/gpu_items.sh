#!/bin/bash
# A/B of the work-item pairing (partners in order of length vs neighbours in file order) on the shapes it matters for.
for w in c4 c5 c2; do
  for o in length file; do
    echo "== $w items in $o order"
    PAIRALIGN_ITEM_ORDER=$o timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --shapes none --no-cli-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','gcups','parity_spot_check')}), d['roofline']['frac'], d['e2e']['gcups'])"
  done
done
