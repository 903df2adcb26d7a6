#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the default bench line under torchrun, then the command line's in-process
# multi-device path on config 3 (-j -n -m) and config 5 (-a -n) with its phase timings.
N=${1:-8}
mkdir -p gpurun_out
TAG=${TAG:-r02}
{
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-3} --warmup 2 2>&1 | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu.json
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read())
print("main", d["n_gpus"], d["scaling"], "value", d["value"], "gcups", d["gcups"], "e2e", d["e2e"]["value"], "ok", d["parity_spot_check"], "cli", d.get("cli_multi_device_md5_equal"))
print("cells/rank", d["config"]["cells_per_rank"], "kernel ms/rank", d["config"]["kernel_ms_per_rank"])
for k, v in d.get("shapes", {}).items():
    print(k, "ms", round(v["ms_per_step"], 1), "gcups", round(v["gcups"]), "e2e gcups", round(v["e2e"]["gcups"]), "ok", v["parity_spot_check_all_ranks"], "frac", round(v["roofline"]["frac"], 3))
PY
python - <<PY
from phylommand_b200 import synth
names, seqs = synth.make_16s_like(10000, 1003)
synth.write_fasta("/tmp/c3.fst", names, seqs)
names, seqs = synth.make_long(200, 1005)
synth.write_fasta("/tmp/c5.fst", names, seqs)
PY
echo "== command line, config 3, -j -n -m, $N devices"
( time PAIRALIGN_TIMING=1 build/pairalign_b200 -j -n -m /tmp/c3.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -12
echo "== command line, config 5, -a -n, $N devices (one device: 9fdbc979ac14b6ca303f83d01628d930)"
( time PAIRALIGN_TIMING=1 build/pairalign_b200 -a -n /tmp/c5.fst | md5sum ) 2>&1 | grep -a -v "^$" | tail -12
} 2>&1 | tee gpurun_out/${TAG}_multi_${N}gpu.log
