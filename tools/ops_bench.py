#!/usr/bin/env python3
"""pairalign -a at BASELINE.json config 5 sizes (200 x 30 kb): DP with stored moves + walk back, through
pa_align_pairs_ops.  Prints one JSON line per run.

    python tools/ops_bench.py [--seqs 200] [--pairs 592] [--length 30000] [--devices 0]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from phylommand_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=200)
    ap.add_argument("--pairs", type=int, default=592)
    ap.add_argument("--length", type=int, default=30000)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--tag", default="")
    ap.add_argument("--waves", type=int, default=0, help="pair list of exactly WAVES x 148 two-pair work items (rows of 148 pairs) instead of a triangle prefix")
    a = ap.parse_args()
    _, seqs = synth.make_long(a.seqs, 1005, length=a.length)
    enc = [synth.to_masks(s) for s in seqs]
    lens = np.array([len(e) for e in enc])
    capi.init([int(d) for d in a.devices.split(",")])
    try:
        capi.upload(enc)
        n = len(enc)
        ia = np.array([x for x in range(n) for y in range(x + 1, n)], dtype=np.uint32)[:a.pairs]
        ib = np.array([y for x in range(n) for y in range(x + 1, n)], dtype=np.uint32)[:a.pairs]
        if a.waves:
            ia = np.repeat(np.arange(2 * a.waves, dtype=np.uint32), 148)
            ib = (ia + 1 + np.tile(np.arange(148, dtype=np.uint32), 2 * a.waves)) % n
        capi.align_pairs_ops(ia[:8], ib[:8], lens)                 # warm-up: allocations, module load
        t0 = time.perf_counter()
        ops, off, n_ops, res = capi.align_pairs_ops(ia, ib, lens)
        dt = time.perf_counter() - t0
        t = capi.timing()
        cells = int((lens[ia].astype(np.int64) * lens[ib]).sum())
        print(json.dumps({"what": "pa_align_pairs_ops", "tag": a.tag, "pairs": len(ia), "cells": cells, "devices": a.devices,
                          "call_s": dt, "gcups_call": cells / dt / 1e9, "pairs_per_s": len(ia) / dt,
                          "kernel_ms": t["kernel_ms"], "walk_ms": t["walk_ms"], "dp_cta_ms": t["dp_cta_ms"], "dp_duo_ms": t["dp_duo_ms"],
                          "dp_general_ms": t["dp_general_ms"],
                          "gcups_dp": cells / max(t["kernel_ms"] - t["walk_ms"], 1e-9) / 1e6,
                          "launches": t["kernel_launches"], "op_bytes": int(n_ops.sum())}))
    finally:
        capi.shutdown()


if __name__ == "__main__":
    main()
