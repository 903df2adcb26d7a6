#!/usr/bin/env python3
"""bench.py -- all-pairs pairalign throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5|tiny]

One "step" = one pass of the hot path over one batch: the all-pairs alignment of the
synthetic set (upper triangle, JC-distance inputs: score, mismatches, compared columns per
pair).  At N=1 the workload is BASELINE.json configs[1]: 1,000 x 1.5 kb 16S-like sequences
(499,500 pairs, ~1.12e12 DP cells).  At N>1 (torchrun, one rank per GPU) the set grows to
round(1000*sqrt(N)) sequences so that every GPU keeps the same number of DP cells (weak
scaling); the triangle is cut into N contiguous ranges balanced by DP cells
(pa_partition_pairs) and there is no collective on the data path.

Keys of the JSON line:
  value      pairs/s with the sequences resident in HBM and the records left in HBM
  e2e        pairs/s through the C-ABI call a user makes (pa_upload_sequences + pa_align_all_pairs)
             with HOST buffers: host packing, H2D, kernels, D2H into a numpy array, every step
  roofline   the DP kernel against the INT32 issue rate measured on this GPU in this run
  cpu_baseline  the reference's own pairalign (oracle/_ref, -O2 -DPTHREAD) on the host cores,
             on a bounded prefix of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
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


def workload(name: str, n_gpus: int, strong: bool = False):
    from phylommand_b200 import synth
    if strong:                      # fixed set: the triangle is cut into n_gpus ranges
        n_gpus = 1
    if name == "c1":
        # BASELINE.json configs[0]: the reference's example file, gap characters removed (the DP path); 103 sequences,
        # 12 of them with IUPAC codes.  Fixed size (no weak scaling); the CPU baseline runs the whole file.
        import numpy as _np
        names, seqs, cur = [], [], None
        for line in (ROOT / "tests" / "golden" / "example_files" / "alignment_file_degapped.fst").read_text().split("\n"):
            if line.startswith(">"):
                names.append(line[1:].split("|")[0].replace(" ", "")); seqs.append("")
            elif names:
                seqs[-1] += line.strip()
        order = sorted(range(len(names)), key=lambda k: names[k].encode())     # the reference's std::map order
        names = [names[k] for k in order]
        seqs = [_np.frombuffer(seqs[k][1:].upper().encode(), dtype=_np.uint8) for k in order]   # first character dropped
        label = f"example_files/alignment_file.fst without gap characters ({len(seqs)} sequences), all-pairs"
    elif name == "c2":
        n = int(round(1000 * (n_gpus ** 0.5)))
        names, seqs = synth.make_16s_like(n, 1002)
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1002), all-pairs"
    elif name == "c3":
        n = int(round(10000 * (n_gpus ** 0.5)))
        names, seqs = synth.make_16s_like(n, 1003)
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1003), all-pairs"
    elif name == "c4":
        n = int(round(5000 * (n_gpus ** 0.5)))
        names, seqs, _ = synth.make_its_like(n, 1004)
        label = f"synthetic {n} ITS-like 400-900 bp (seed 1004), all-pairs"
    elif name == "c5":
        n = int(round(200 * (n_gpus ** 0.5)))
        names, seqs = synth.make_long(n, 1005)
        label = f"synthetic {n} x 30 kb (seed 1005), all-pairs"
    elif name == "c5s":
        n = int(round(16 * (n_gpus ** 0.5)))
        names, seqs = synth.make_long(n, 1005)
        label = f"synthetic {n} x 30 kb (seed 1005), all-pairs [profiling size]"
    elif name == "c2n":
        # config 2 with two IUPAC ambiguity codes in every sequence: every pair takes the general (4-bit) kernel
        import numpy as _np
        n = int(round(1000 * (n_gpus ** 0.5)))
        names, seqs = synth.make_16s_like(n, 1002)
        rng = _np.random.default_rng(5)
        amb = _np.frombuffer(b"RYSWKMN", dtype=_np.uint8)
        seqs = [s.copy() for s in seqs]
        for s_ in seqs:
            s_[rng.integers(1, len(s_), size=2)] = amb[rng.integers(0, len(amb), size=2)]
        label = f"synthetic {n} x 1.5 kb 16S-like (seed 1002) with 2 IUPAC codes per sequence, all-pairs"
    elif name == "c5w":
        n = int(round(64 * (n_gpus ** 0.5)))
        names, seqs = synth.make_long(n, 1005, length=7600, spread=0.05)
        label = f"synthetic {n} x 7.6 kb (seed 1005), all-pairs [profiling size of the floating-window s16x2 path]"
    elif name == "tiny":
        n = int(round(128 * (n_gpus ** 0.5)))
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


def cpu_prefix_for(threads: int) -> int:
    """Prefix size giving roughly 10 s of reference work: ~0.23 core-seconds per 1.5 kb pair at -O2."""
    return int(min(160, max(16, round(9.0 * threads ** 0.5))))


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
    k = args.cpu_prefix or (len(seqs) if args.workload == "c1" else cpu_prefix_for(threads))
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


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c2n", "c3", "c4", "c5", "c5s", "c5w", "tiny"])
    ap.add_argument("--cpu-prefix", type=int, default=0, help="sequences in the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peak", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): the set grows with sqrt(N) so every GPU keeps the cells of one; strong: fixed set")
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

    names, seqs, label = workload(args.workload, world, args.scaling == "strong")
    enc = [synth.to_masks(s) for s in seqs]
    masks, offsets = capi.pack(enc)
    capi.upload_packed(masks, offsets)
    total_pairs = capi.num_pairs()
    bounds = capi.partition_pairs(0, total_pairs, world)
    first, count = int(bounds[rank]), int(bounds[rank + 1] - bounds[rank])
    my_cells = capi.count_cells(first, count)
    total_cells = capi.count_cells(0, total_pairs)

    d_out = torch.empty(count * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    h_out = np.empty(count, dtype=capi.RESULT_DTYPE)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        capi.align_all_pairs_device(d_out.data_ptr(), first, count)
        return capi.timing()

    e2e_parts = {"upload_ms": 0.0, "align_ms": 0.0, "kernel_ms": 0.0, "d2h_ms": 0.0, "calls": 0}

    def step_e2e():
        t0 = time.perf_counter()
        capi.upload_packed(masks, offsets)
        t1 = time.perf_counter()
        capi.align_all_pairs(first, count, h_out)
        t = capi.timing()
        e2e_parts["upload_ms"] += 1e3 * (t1 - t0)
        e2e_parts["align_ms"] += t["total_ms"]
        e2e_parts["kernel_ms"] += t["kernel_ms"]
        e2e_parts["d2h_ms"] += t["d2h_ms"]
        e2e_parts["calls"] += 1
        return t

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        el, kern, launches = 0.0, 0.0, 0
        for _ in range(steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            t = fn()
            torch.cuda.synchronize()
            el += time.perf_counter() - t0
            kern += t["dp_duo_ms"] + t["dp_fast_ms"] + t["dp_cta_ms"] + t["dp_general_ms"]
            launches += t["kernel_launches"]
        barrier()
        v = torch.tensor([el, kern], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v[0]), float(v[1]), launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    el, kern_ms, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    el_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 3))

    # parity spot check of what was just timed (a sample of this rank's records against the oracle)
    ok = None
    if rank == 0:
        from tests import oracle_lib
        oracle = oracle_lib.load()
        rng = np.random.default_rng(0)
        ok = True
        for q in rng.choice(count, size=min(8, count), replace=False):
            a, b = capi.pair_from_index(first + int(q))
            ok = ok and tuple(h_out[int(q)]) == tuple(oracle.align_forward(enc[a], enc[b]))
        dev = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
        ok = ok and dev.tobytes() == h_out.tobytes()

    peak = None
    if rank == 0 and not args.no_peak:
        peak = {capi.PEAK_CLASSES[w]: capi.int32_peak(w) for w in (0, 12, 2, 3, 7, 8, 9, 18, 21)}

    if rank == 0:
        ms = 1e3 * el / args.steps
        value = total_pairs / (el / args.steps)
        kern_step_ms = kern_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int32",
            "data": "reference example file" if args.workload == "c1" else "synthetic", "gcups": total_cells / (el / args.steps) / 1e9,
            "config": {"workload": label, "mode": "-j -m (Jukes-Cantor matrix inputs: score, mismatches, columns per pair)",
                       "pairs": total_pairs, "cells": total_cells, "scoring": "match 7 / mismatch -5 / gap open -15 / extend -1",
                       "sharding": f"triangle cut into {world} contiguous ranges balanced by DP cells, no collective",
                       "l2": "256 MiB device write between steps, outside the timed region; the kernel is ALU-bound"},
            "e2e": {"value": total_pairs / (el_e2e / args.steps), "unit": "pairs/s", "h2d_bytes_per_step": int(masks.nbytes + offsets.nbytes),
                    "d2h_bytes_per_step": int(count * capi.RESULT_DTYPE.itemsize), "ms_per_step": 1e3 * el_e2e / args.steps,
                    "gcups": total_cells / (el_e2e / args.steps) / 1e9,
                    "rank0_breakdown_ms_per_call": {k: v / max(1, e2e_parts["calls"]) for k, v in e2e_parts.items() if k != "calls"},
                    "path": "pa_upload_sequences(host masks) + pa_align_all_pairs(host records): pack, H2D, kernels, D2H, copy-out"},
            "gpu_launches": launches, "clocks": clocks, "parity_spot_check": ok,
        }
        if peak is not None:
            # Single-instruction issue rates measured on this GPU a moment ago (pa_int32_peak): every integer
            # class the DP uses -- 32-bit or s16x2 -- issues at the same ~64 lanes/clk/SM.  The s16x2 DPX forms
            # carry two DP cells per lane, so the roofline of the packed kernel is twice the 32-bit-lane rate.
            alu32 = max(peak["IADD3"][0], peak["VIMNMX"][0], peak["VIADDMNMX"][0])
            packed = max(peak["VIMNMX3.S16x2"][0], peak["VIADDMNMX.S16x2"][0])
            peak_ops = 2.0 * packed
            achieved = OPS_PER_CELL * my_cells / (kern_step_ms * 1e-3) / 1e9
            line["roofline"] = {
                "bound": "int32", "achieved": achieved, "peak": peak_ops, "unit": "Gop/s", "frac": achieved / peak_ops,
                "traffic": 7372288 if (args.workload == "c2" and world == 1) else None,
                "kernel": "pa_warp_duo_kernel<0,1,-1> (s16x2 DPX, two pairs per warp, two rows per step, strip width 8-13 columns per lane chosen per work item, biased storage with H+GO on the FMA pipe)",
                "kernel_ms_per_step": kern_step_ms, "kernel_gcups": my_cells / (kern_step_ms * 1e-3) / 1e9,
                "ops_per_cell": OPS_PER_CELL,
                "achieved_def": "14 integer operations per DP cell (SURVEY.md 8d) x cells of this rank / CUDA-event time of the DP kernels",
                "peak_def": "2 x the measured issue rate of VIMNMX3.S16x2 / VIADDMNMX.S16x2 chains (two 16-bit cells per 32-bit lane), "
                            "all SMs, measured in this run by pa_int32_peak",
                "frac_vs_32bit_lane_roofline": achieved / alu32, "peak_32bit_lane": alu32,
                "traffic_def": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                               "(profiles/r01_v5_duo_bias_ncu_full.txt: 0.99 MB read + 6.39 MB written); algorithmic bytes per launch are in hbm.algorithmic_bytes_per_step",
                "measured": {k: {"gops": v[0], "sm_mhz": v[1]} for k, v in peak.items()},
                "hbm": {"algorithmic_bytes_per_step": int(masks.nbytes // 4 + count * 20),
                        "note": "2-bit sequences + 20 B per pair; HBM is not the bound (SURVEY.md 8d)"},
            }
        if not args.no_cpu_baseline:
            threads = min(host_threads(), 64)
            k = args.cpu_prefix or (len(seqs) if args.workload == "c1" else cpu_prefix_for(threads))
            cb = cpu_reference_run(names, seqs, k, threads)
            line["cpu_baseline"] = {"value": cb["pairs"] / cb["seconds"], "unit": "pairs/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"], "gcups": cb["cells"] / cb["seconds"] / 1e9, "seconds": cb["seconds"]}
        print(json.dumps(line), flush=True)
    capi.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
