#!/usr/bin/env python3
"""bench.py -- all-pairs pairalign throughput on B200 (BASELINE.json metric: pairs/s and GCUPS at 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c3s|c2|c3|c4|c5|...] [--scaling strong|weak] [--shapes auto|none|c2,c4,...]

One "step" = one pass of the hot path over one batch: the all-pairs alignment of the synthetic set (upper triangle,
JC-distance inputs: score, mismatches, compared columns per pair).

Default line: STRONG scaling on a FIXED set -- the first 4,000 sequences of BASELINE.json configs[2] (10,000 x 1.5 kb,
seed 1003): 7,998,000 pairs, 1.8e13 DP cells, the same set at every N.  The triangle is cut into N contiguous ranges
balanced by DP cells (pa_partition_pairs), one rank per GPU under torchrun, no collective on the data path; every rank
uploads the whole set (6 MB) and copies its own range of records back.  The driver's scaling efficiency therefore
measures partition balance, the tail wave and the per-rank fixed costs (upload, D2H, copy-out).
`--scaling weak` keeps round 1's line (set grows with sqrt(N)).

`shapes` (same JSON line) holds one short measurement of every other BASELINE.json configuration at this N, same
machinery (strong scaling): c2 = configs[1] (1,000 x 1.5 kb), c3 = configs[2] in full (10,000 x 1.5 kb, the
north_star target), c4 = configs[3] (5,000 x 400-900 bp), c5 = configs[4] (200 x 30 kb).

Keys of the JSON line:
  value      pairs/s with the sequences resident in HBM and the records left in HBM (max over ranks of the step time)
  e2e        pairs/s through the C-ABI call a user makes (pa_upload_sequences + pa_align_all_pairs) with HOST
             buffers: H2D of the raw sets, packing kernel, DP kernels, D2H into a numpy array, every step
  roofline   the DP kernels against the INT32 issue rate measured on this GPU in this run (pa_int32_peak)
  cpu_baseline  the reference's own pairalign (oracle/_ref, -O2 -DPTHREAD) on the host cores, on a bounded prefix
             of the same workload, median of 3 warmed runs
  cli_multi_device_md5_equal (N > 1)  the in-process multi-device path of the command line (one process, one host
             thread per GPU) prints the same bytes on N devices as on one (256-sequence prefix)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

OPS_PER_CELL = 14          # INT32-pipe operations per DP cell in stats mode
METRIC = "pairs/sec (all-pairs pairalign: seqpair DP + per-pair distance statistics)"

#: the DP kernel behind each pa_timing bucket
KERNELS = {
    "dp_duo_ms": "pa_warp_duo_kernel (s16x2 DPX, two pairs per warp, two rows per step, strip width 8-13 columns per "
                 "lane per work item; floating 16-bit window for pairs beyond int16) / pa_warp_sets_kernel (the same recurrence on 4-bit IUPAC sets) when sequences carry ambiguity codes",
    "dp_fast_ms": "pa_warp32_kernel<16> (int32, one pair per warp)",
    "dp_cta_ms": "pa_cta_duo_moves_kernel (s16x2 with stored moves, one two-pair item per CTA, shared-memory ring edges; statistics "
                 "by the walk) or pa_cta32_kernel<16> (int32, one pair per CTA) for long pairs mixed with short ones",
    "dp_general_ms": "pa_warp_dp_kernel<8,true> (int32 with wrap-around, 4-bit IUPAC sets)",
    "walk_ms": "pa_walk_kernel (statistics of few long pairs: walk over the stored moves of pa_cta_duo_moves_kernel)",
}
#: ncu --set full summaries under profiles/ that belong to a workload's dominant kernel (traffic and pipe figures are
#: only reported when the file for THIS workload exists); NCU_CELLS: DP cells of the launch a summary was captured on
NCU_FILES = {"c2": ["r02_duo_ncu_full.txt", "r01_v5_duo_bias_ncu_full.txt"], "c3s": ["r02_duo_ncu_full.txt", "r01_v5_duo_bias_ncu_full.txt"],
             "c3": ["r02_duo_ncu_full.txt", "r01_v5_duo_bias_ncu_full.txt"], "c5w": ["r01_v4_duo_win_ncu_full.txt"],
             "c2n": ["r02_sets_ncu_full.txt"]}

NCU_CELLS = {"r02_duo_ncu_full.txt": 1123146785408, "r02_sets_ncu_full.txt": 1123146785408}     # both: config 2, one launch

_SETS: dict = {}


def _c3_full():
    from phylommand_b200 import synth
    if "c3" not in _SETS:
        _SETS["c3"] = synth.make_16s_like(10000, 1003)
    return _SETS["c3"]


def workload(name: str, n_gpus: int, strong: bool = True):
    """(names, sequences, label).  Weak scaling multiplies the sequence count by sqrt(N); strong keeps the set."""
    from phylommand_b200 import synth
    g = 1.0 if strong else n_gpus ** 0.5
    if name == "c1":
        # BASELINE.json configs[0]: the reference's example file, gap characters removed (the DP path); 103 sequences,
        # 12 of them with IUPAC codes.  Fixed size; the CPU baseline runs the whole file.
        names, seqs = [], []
        for line in (ROOT / "tests" / "golden" / "example_files" / "alignment_file_degapped.fst").read_text().split("\n"):
            if line.startswith(">"):
                names.append(line[1:].split("|")[0].replace(" ", "")); seqs.append("")
            elif names:
                seqs[-1] += line.strip()
        order = sorted(range(len(names)), key=lambda k: names[k].encode())     # the reference's std::map order
        names = [names[k] for k in order]
        seqs = [np.frombuffer(seqs[k][1:].upper().encode(), dtype=np.uint8) for k in order]   # first character dropped
        label = f"example_files/alignment_file.fst without gap characters ({len(seqs)} sequences), all-pairs"
    elif name == "c3s":
        names, seqs = _c3_full()
        n = min(len(seqs), int(round(4000 * g)))
        names, seqs = names[:n], seqs[:n]
        label = (f"first {n} sequences of BASELINE configs[2] (synthetic 10,000 x 1.5 kb 16S-like, seed 1003), all-pairs; "
                 "the full set is under shapes.c3")
    elif name == "c2":
        n = int(round(1000 * g))
        names, seqs = synth.make_16s_like(n, 1002)
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1002), all-pairs"
    elif name == "c3":
        if strong:
            names, seqs = _c3_full()
        else:
            names, seqs = synth.make_16s_like(int(round(10000 * g)), 1003)
        label = f"synthetic {len(seqs)} x 1.5 kb 16S-like (seed 1003), all-pairs"
    elif name == "c4":
        n = int(round(5000 * g))
        names, seqs, _ = synth.make_its_like(n, 1004)
        label = f"synthetic {n} ITS-like 400-900 bp (seed 1004), all-pairs"
    elif name == "c5":
        n = int(round(200 * g))
        names, seqs = synth.make_long(n, 1005)
        label = f"synthetic {n} x 30 kb (seed 1005), all-pairs"
    elif name == "c5s":
        n = int(round(16 * g))
        names, seqs = synth.make_long(n, 1005)
        label = f"synthetic {n} x 30 kb (seed 1005), all-pairs [profiling size]"
    elif name == "c2n":
        # config 2 with two IUPAC ambiguity codes in every sequence
        n = int(round(1000 * g))
        names, seqs = synth.make_16s_like(n, 1002)
        rng = np.random.default_rng(5)
        amb = np.frombuffer(b"RYSWKMN", dtype=np.uint8)
        seqs = [s.copy() for s in seqs]
        for s_ in seqs:
            s_[rng.integers(1, len(s_), size=2)] = amb[rng.integers(0, len(amb), size=2)]
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1002) with 2 IUPAC codes per sequence, all-pairs"
    elif name == "c5w":
        n = int(round(64 * g))
        names, seqs = synth.make_long(n, 1005, length=7600, spread=0.05)
        label = f"synthetic {n} x 7.6 kb (seed 1005), all-pairs [profiling size of the floating-window s16x2 path]"
    elif name == "tiny":
        n = int(round(128 * g))
        names, seqs = synth.make_16s_like(n, 1002)
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1002), all-pairs [tiny]"
    else:
        raise SystemExit(f"unknown workload {name}")
    return names, seqs, label


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(names, seqs, k_prefix: int, threads: int):
    """Time the reference's own pairalign (-j -n -m) on the first k sequences.  Returns dict."""
    from phylommand_b200 import synth
    from tests import oracle_lib
    k = min(k_prefix, len(seqs))
    pairs = k * (k - 1) // 2
    lens = np.array([len(s) for s in seqs[:k]], dtype=np.int64)
    cells = int((lens.sum() ** 2 - (lens ** 2).sum()) // 2)
    exe = oracle_lib.REF_CLI_PTHREAD if oracle_lib.REF_CLI_PTHREAD.exists() else None
    if exe is not None:
        with tempfile.TemporaryDirectory() as td:
            fa = Path(td) / "prefix.fst"
            synth.write_fasta(fa, names[:k], seqs[:k])
            t0 = time.perf_counter()
            r = subprocess.run([str(exe), "-T", str(threads), "-j", "-n", "-m", str(fa)], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
        if r.returncode == 0:
            return dict(kind="reference", seconds=dt, pairs=pairs, cells=cells, cores=threads,
                        sample=f"oracle/_ref/pairalign_pthread -T {threads} -j -n -m on the first {k} sequences "
                               f"({pairs} pairs, {cells:.3e} cells)")
    oracle = oracle_lib.load()
    enc = [synth.to_masks(s) for s in seqs[:k]]
    masks = np.concatenate(enc)
    offsets = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.uint64)
    t0 = time.perf_counter()
    oracle.all_pairs(masks, offsets, threads=threads)
    dt = time.perf_counter() - t0
    return dict(kind="port", seconds=dt, pairs=pairs, cells=cells, cores=threads,
                sample=f"oracle/pa_oracle.c forward port, {threads} threads, first {k} sequences ({pairs} pairs, {cells:.3e} cells)")


def cpu_prefix_for(threads: int, seqs) -> int:
    """Prefix size giving roughly 5 s of reference work: ~0.1 core-microseconds per DP cell at -O2."""
    mean = float(np.mean([len(s) for s in seqs[:64]]))
    k = int(round((2.0 * 5.0 * threads / (1.0e-7 * mean * mean)) ** 0.5))
    return int(min(len(seqs), max(8, k)))


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    names, seqs, label = workload(args.workload, args.gpus, args.scaling == "strong")
    threads = min(host_threads(), 64)
    k = args.cpu_prefix or (len(seqs) if args.workload == "c1" else cpu_prefix_for(threads, seqs))
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        res = cpu_reference_run(names, seqs, k, threads)
        if it >= args.warmup:
            times.append(res["seconds"])
    ms = 1e3 * sum(times) / len(times)
    value = res["pairs"] / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "gcups": res["cells"] / (ms * 1e-3) / 1e9,
            "config": {"workload": label, "mode": "-j -n -m (Jukes-Cantor matrix)", "step": res["sample"]},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ncu_summary_for(workload_name: str):
    """Pipe utilisation and DRAM traffic of the workload's dominant kernel from a committed ncu --set full summary
    (profiles/), or None when no capture of THIS workload's kernel exists."""
    for fn in NCU_FILES.get(workload_name, []):
        p = ROOT / "profiles" / fn
        if not p.exists():
            continue
        vals = {}
        for ln in p.read_text().splitlines():        # tools/ncu_summary.py lines: "<metric>  <value> <unit>"
            m = re.match(r"^(\S+)\s+([-\d.,]+)\s*(\S*)\s*$", ln)
            if m:
                scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1.0)
                vals.setdefault(m.group(1), float(m.group(2).replace(",", "")) * scale)
        out = {"file": f"profiles/{fn}",
               "alu_pipe_pct": vals.get("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
               "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "dram_bytes": None, "warp_inst": vals.get("smsp__inst_executed.sum"), "inst_per_cell": None}
        lanes = vals.get("smsp__thread_inst_executed_per_inst_executed.ratio")
        if out["warp_inst"] and lanes and fn in NCU_CELLS:       # thread instructions executed per DP cell, everything included
            out["inst_per_cell"] = out["warp_inst"] * lanes / NCU_CELLS[fn]
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            out["dram_bytes"] = int(rd + wr)
        return out
    return None


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3s", choices=["c1", "c2", "c2n", "c3", "c3s", "c4", "c5", "c5s", "c5w", "tiny"])
    ap.add_argument("--cpu-prefix", type=int, default=0, help="sequences in the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peak", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the same fixed set at every N; weak: the set grows with sqrt(N)")
    ap.add_argument("--shapes", default="auto", help="auto: c2,c4,c5,c3 when the main workload is c3s; none; or a list")
    ap.add_argument("--no-cli-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world

    import torch
    import torch.distributed as dist
    from phylommand_b200 import build, capi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    build.build_library()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.init([local_rank])
    strong = args.scaling == "strong"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        v = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return [float(x) for x in v]

    def gather_all(val: float):
        v = torch.tensor([val], dtype=torch.float64, device="cuda")
        if world > 1:
            out = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(out, v)
            return [float(x[0]) for x in out]
        return [val]

    oracle = None

    def measure(name: str, steps: int, warmup: int, e2e_steps: int, e2e_warmup: int, sample_clocks: bool = False):
        """One workload at this N: resident kernel-path timing, end-to-end timing, per-rank parity spot check."""
        nonlocal oracle
        names, seqs, label = workload(name, world, strong)
        enc = [synth.to_masks(s) for s in seqs]
        masks, offsets = capi.pack(enc)
        capi.upload_packed(masks, offsets)
        total_pairs = capi.num_pairs()
        bounds = capi.partition_pairs(0, total_pairs, world)
        first, count = int(bounds[rank]), int(bounds[rank + 1] - bounds[rank])
        my_cells = capi.count_cells(first, count)
        total_cells = capi.count_cells(0, total_pairs)
        d_out = torch.empty(max(count, 1) * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        h_out = np.empty(count, dtype=capi.RESULT_DTYPE)
        h_out.view(np.uint8).reshape(-1)[:] = 0               # touched (np.zeros would not): first-touch page faults are not part of a step
        parts = {"upload_ms": 0.0, "align_ms": 0.0, "kernel_ms": 0.0, "d2h_ms": 0.0, "calls": 0}
        buckets = {k: 0.0 for k in KERNELS}

        def step_resident():
            capi.align_all_pairs_device(d_out.data_ptr(), first, count)
            return capi.timing()

        def step_e2e():
            t0 = time.perf_counter()
            capi.upload_packed(masks, offsets)
            t1 = time.perf_counter()
            capi.align_all_pairs(first, count, h_out)
            t = capi.timing()
            parts["upload_ms"] += 1e3 * (t1 - t0); parts["align_ms"] += t["total_ms"]
            parts["kernel_ms"] += t["kernel_ms"]; parts["d2h_ms"] += t["d2h_ms"]; parts["calls"] += 1
            return t

        def timed(fn, n_steps, n_warm, acc=None):
            for _ in range(n_warm):
                fn()
            el, kern, launches = 0.0, 0.0, 0
            for _ in range(n_steps):
                flush.zero_()
                barrier()
                t0 = time.perf_counter()
                t = fn()
                torch.cuda.synchronize()
                el += time.perf_counter() - t0
                kern += sum(t[k] for k in KERNELS)
                launches += t["kernel_launches"]
                if acc is not None:
                    for k in KERNELS:
                        acc[k] += t[k]
            barrier()
            el_max, kern_max = reduce_max([el, kern])
            return el_max, kern_max, launches

        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        el, kern_ms, launches = timed(step_resident, steps, warmup, buckets)
        clocks = sampler.stop() if sampler else None
        if e2e_steps is None:
            # the K of the resident timing, capped so that the end-to-end leg of a long step (7 s on the default set at
            # N = 1) stays within about a minute; el is the max over ranks, so every rank takes the same number
            per = el / steps
            e2e_steps = steps if per * steps <= 60.0 else min(steps, max(3, int(60.0 / per)))
        el_e2e, _, _ = timed(step_e2e, e2e_steps, e2e_warmup)

        # parity spot check of what was just timed: EVERY rank checks a sample of its own records against the oracle
        # (and its device-resident records against the host ones); the line carries the AND over ranks
        from tests import oracle_lib
        if oracle is None:
            oracle = oracle_lib.load()
        ok = True
        lens = np.array([len(e) for e in enc], dtype=np.int64)
        rng = np.random.default_rng(rank)
        n_check = 8 if int(lens.max()) <= 4096 else 2            # a 30 kb pair costs the oracle ~10 s
        for q in rng.choice(count, size=min(n_check, count), replace=False) if count else []:
            a, b = capi.pair_from_index(first + int(q))
            ok = ok and tuple(h_out[int(q)]) == tuple(oracle.align_forward(enc[a], enc[b]))
        dev = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)[:count]
        ok = ok and dev.tobytes() == h_out.tobytes()
        ok_all = all(v > 0.5 for v in gather_all(1.0 if ok else 0.0))
        cells_rank = gather_all(float(my_cells))
        kern_rank = gather_all(sum(buckets.values()))
        del d_out
        res = dict(names=names, seqs=seqs, label=label, total_pairs=total_pairs, total_cells=total_cells, my_cells=my_cells,
                   count=count, masks_bytes=int(masks.nbytes + offsets.nbytes), el=el / steps, kern_ms=kern_ms / steps,
                   launches=launches, el_e2e=el_e2e / e2e_steps, e2e_steps=e2e_steps, parts=parts, buckets=buckets, clocks=clocks, ok=ok_all,
                   cells_rank=cells_rank, kern_rank=[k / steps for k in kern_rank], max_len=int(lens.max()))
        return res

    main_res = measure(args.workload, args.steps, args.warmup, None, max(1, args.warmup // 3), sample_clocks=True)

    peak = None
    if not args.no_peak:
        # every rank measures (keeps the ranks in step); rank 0's figures are reported
        peak = {capi.PEAK_CLASSES[w]: capi.int32_peak(w) for w in (0, 12, 2, 3, 7, 8, 9, 18, 21)}
    peak_ops = alu32 = None
    if peak is not None:
        # Single-instruction issue rates measured on this GPU a moment ago (pa_int32_peak): every integer class the DP
        # uses -- 32-bit or s16x2 -- issues at the same ~64 lanes/clk/SM.  The s16x2 DPX forms carry two DP cells per
        # lane, so the roofline of the packed kernel is twice the 32-bit-lane rate.
        alu32 = max(peak["IADD3"][0], peak["VIMNMX"][0], peak["VIADDMNMX"][0])
        peak_ops = 2.0 * max(peak["VIMNMX3.S16x2"][0], peak["VIADDMNMX.S16x2"][0])

    def roofline_of(r, name):
        dom = max(r["buckets"], key=lambda k: r["buckets"][k])
        # the CTA bucket holds the s16x2 move kernel when the walk ran (few long items), the int32 CTA kernel otherwise
        packed = dom == "dp_duo_ms" or (dom == "dp_cta_ms" and r["buckets"]["walk_ms"] > 0)
        pk = peak_ops if packed else alu32
        # max over ranks of the kernel time against that rank's cells: use the slowest rank's figure (whole-job view)
        achieved = OPS_PER_CELL * (r["total_cells"] / world) / (r["kern_ms"] * 1e-3) / 1e9
        ncu = ncu_summary_for(name)
        out = {"bound": "int32", "achieved": achieved, "peak": pk, "unit": "Gop/s", "frac": achieved / pk if pk else None,
               "traffic": ncu["dram_bytes"] if ncu else None, "kernel": KERNELS[dom],
               "kernel_share_of_dp_time": r["buckets"][dom] / max(1e-9, sum(r["buckets"].values())),
               "kernel_ms_per_step": r["kern_ms"], "kernel_gcups_per_gpu": (r["total_cells"] / world) / (r["kern_ms"] * 1e-3) / 1e9,
               "ops_per_cell": OPS_PER_CELL}
        if ncu:
            out["ncu"] = {"file": ncu["file"], "alu_pipe_pct": ncu["alu_pipe_pct"], "issue_active_pct": ncu["issue_active_pct"],
                          "inst_per_cell": ncu["inst_per_cell"],
                          "note": "from the committed ncu --set full capture of this kernel on this workload's pair shape"}
        return out

    def shape_entry(r, name):
        e = {"workload": r["label"], "pairs": r["total_pairs"], "cells": r["total_cells"], "ms_per_step": 1e3 * r["el"],
             "value": r["total_pairs"] / r["el"], "unit": "pairs/s", "gcups": r["total_cells"] / r["el"] / 1e9,
             "e2e": {"value": r["total_pairs"] / r["el_e2e"], "ms_per_step": 1e3 * r["el_e2e"],
                     "gcups": r["total_cells"] / r["el_e2e"] / 1e9},
             "parity_spot_check_all_ranks": r["ok"],
             "cells_per_rank": {"min": min(r["cells_rank"]), "max": max(r["cells_rank"])},
             "kernel_ms_per_rank": {"min": min(r["kern_rank"]), "max": max(r["kern_rank"])}}
        if peak is not None:
            rf = roofline_of(r, name)
            e["roofline"] = {k: rf[k] for k in ("achieved", "peak", "frac", "kernel", "kernel_ms_per_step", "kernel_gcups_per_gpu")}
        return e

    shapes = {}
    shape_list = []
    if args.shapes == "auto":
        shape_list = ["c2", "c4", "c5", "c3"] if args.workload == "c3s" else []
    elif args.shapes != "none":
        shape_list = [s for s in args.shapes.split(",") if s]
    for nm in shape_list:
        if nm == "c3":      # the target set in full: 45 s per pass on one GPU
            r = measure("c3", 1 if world == 1 else 2, 0 if world == 1 else 1, 1, 0)
        elif nm == "c5":
            r = measure("c5", 1 if world == 1 else 2, 1, 1, 0)
        else:
            r = measure(nm, 3, 1, 2, 1)
        shapes[nm] = shape_entry(r, nm)

    # the in-process multi-device path of the command line (one process, one host thread per GPU): N devices must
    # print the same bytes as one
    cli_equal = None
    if world > 1 and not args.no_cli_check:
        barrier()
        if rank == 0:
            try:
                exe = build.build_cli()
                names, seqs = main_res["names"][:256], main_res["seqs"][:256]
                with tempfile.TemporaryDirectory() as td:
                    fa = Path(td) / "prefix256.fst"
                    synth.write_fasta(fa, names, seqs)
                    md5 = []
                    for devs in ("0", ",".join(str(k) for k in range(world))):
                        rr = subprocess.run([str(exe), "-j", "-n", "-m", str(fa)], capture_output=True, timeout=600,
                                            env=dict(os.environ, PAIRALIGN_DEVICES=devs))
                        md5.append((rr.returncode, hashlib.md5(rr.stdout).hexdigest(), len(rr.stdout)))
                cli_equal = bool(md5[0] == md5[1] and md5[0][0] == 0 and md5[0][2] > 0)
            except Exception as exc:      # noqa: BLE001 -- reported in the line, never fatal for the bench
                cli_equal = f"error: {exc}"
        barrier()

    if rank == 0:
        r = main_res
        ms = 1e3 * r["el"]
        line = {
            "metric": METRIC, "value": r["total_pairs"] / r["el"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int32", "data": "reference example file" if args.workload == "c1" else "synthetic",
            "gcups": r["total_cells"] / r["el"] / 1e9,
            "config": {"workload": r["label"], "mode": "-j -m (Jukes-Cantor matrix inputs: score, mismatches, columns per pair)",
                       "pairs": r["total_pairs"], "cells": r["total_cells"], "scoring": "match 7 / mismatch -5 / gap open -15 / extend -1",
                       "sharding": f"triangle cut into {world} contiguous ranges balanced by DP cells, no collective",
                       "cells_per_rank": {"min": min(r["cells_rank"]), "max": max(r["cells_rank"])},
                       "kernel_ms_per_rank": {"min": min(r["kern_rank"]), "max": max(r["kern_rank"])},
                       "l2": "256 MiB device write between steps, outside the timed region; the kernel is ALU-bound"},
            "e2e": {"value": r["total_pairs"] / r["el_e2e"], "unit": "pairs/s", "steps": r["e2e_steps"], "h2d_bytes_per_step": r["masks_bytes"],
                    "d2h_bytes_per_step": int(r["count"] * capi.RESULT_DTYPE.itemsize), "ms_per_step": 1e3 * r["el_e2e"],
                    "gcups": r["total_cells"] / r["el_e2e"] / 1e9,
                    "rank0_breakdown_ms_per_call": {k: v / max(1, r["parts"]["calls"]) for k, v in r["parts"].items() if k != "calls"},
                    "path": "pa_upload_sequences(host sets: H2D + packing kernel) + pa_align_all_pairs(host records): kernels, D2H, copy-out"},
            "gpu_launches": r["launches"], "clocks": r["clocks"], "parity_spot_check": r["ok"],
            "parity_spot_check_def": "every rank: 8 of its own records against the oracle and its resident records against the host ones; AND over ranks",
        }
        if peak is not None:
            rf = roofline_of(r, args.workload)
            rf.update({
                "achieved_def": "14 integer operations per DP cell (SURVEY.md 8d) x cells per GPU / CUDA-event time of the DP kernels (max over ranks)",
                "peak_def": "s16x2 kernel: 2 x the measured issue rate of VIMNMX3.S16x2 / VIADDMNMX.S16x2 chains (two 16-bit cells per "
                            "32-bit lane); int32 kernels: the measured IADD3 / VIMNMX / VIADDMNMX rate; all SMs, measured in this run by pa_int32_peak",
                "frac_vs_32bit_lane_roofline": rf["achieved"] / alu32, "peak_32bit_lane": alu32,
                "traffic_def": "dram__bytes_read.sum + dram__bytes_write.sum of one launch from the ncu --set full summary named in "
                               "roofline.ncu.file (null when no capture of this workload's kernel is committed)",
                "measured": {k: {"gops": v[0], "sm_mhz": v[1]} for k, v in peak.items()},
                "hbm": {"algorithmic_bytes_per_step": int(r["masks_bytes"] // 4 + r["count"] * 20),
                        "note": "2-bit sequences + 20 B per pair; HBM is not the bound (SURVEY.md 8d)"}})
            line["roofline"] = rf
        if shapes:
            line["shapes"] = shapes
            if "c3" in shapes and world > 1:      # the north_star target under its own name: the full 10,000 x 1.5 kb set, fixed, cut over N GPUs
                line["c3_strong"] = shapes["c3"]
        if cli_equal is not None:
            line["cli_multi_device_md5_equal"] = cli_equal
        if not args.no_cpu_baseline:
            threads = min(host_threads(), 64)
            k = args.cpu_prefix or (len(r["seqs"]) if args.workload == "c1" else cpu_prefix_for(threads, r["seqs"]))
            runs = [cpu_reference_run(r["names"], r["seqs"], k, threads) for _ in range(4)][1:]     # 1 warm-up + 3
            runs.sort(key=lambda c: c["seconds"])
            cb = runs[1]
            line["cpu_baseline"] = {"value": cb["pairs"] / cb["seconds"], "unit": "pairs/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"] + "; median of 3 runs after 1 warm-up", "gcups": cb["cells"] / cb["seconds"] / 1e9,
                                    "seconds": cb["seconds"], "seconds_all": [c["seconds"] for c in runs]}
        print(json.dumps(line), flush=True)
    capi.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
