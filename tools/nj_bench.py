#!/usr/bin/env python3
"""Neighbour joining timing: pa_nj_build on synthetic tree-like matrices, the achieved
HBM rate against MEASURED_PEAKS.json, and the unmodified reference treeator -n (oracle/_ref) on a bounded size.

    python tools/nj_bench.py [--taxa 1000,2000,5000,10000] [--ref-taxa 1000]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def matrix(n, seed=11):
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 8)).astype(np.float32)
    g = pts @ pts.T
    sq = np.diag(g)
    d = np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2 * g, 0)) + 0.05 * rng.random((n, n), dtype=np.float32)
    return d[np.triu_indices(n, 1)].astype(np.float32)


def write_matrix(path, tri, n):
    with open(path, "w") as fh:
        k = 0
        for a in range(n - 1):
            if a:
                fh.write("\n")
            fh.write(f"t{a:05d} " + " " * a + " ".join("%g" % v for v in tri[k:k + n - 1 - a]) + " ")
            k += n - 1 - a
        fh.write(f"\nt{n - 1:05d}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", default="1000,2000,5000,10000")
    ap.add_argument("--ref-taxa", type=int, default=1000)
    args = ap.parse_args()
    from phylommand_b200 import capi
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs")
    capi.init([0])
    capi.nj_build(matrix(64))                                  # context / module load
    for n in [int(x) for x in args.taxa.split(",")]:
        tri = matrix(n)
        t0 = time.perf_counter()
        res = capi.nj_build(tri)
        wall = time.perf_counter() - t0
        gbs = res["bytes"] / res["kernel_ms"] / 1e6
        print(json.dumps({"taxa": n, "kernel_ms": round(res["kernel_ms"], 3), "call_ms": round(wall * 1e3, 3),
                          "launches": res["launches"], "algorithmic_bytes": res["bytes"], "achieved_gbs": round(gbs, 1),
                          "hbm_peak_gbs": hbm, "frac": round(gbs / hbm, 4) if hbm else None,
                          "us_per_join": round(res["kernel_ms"] * 1e3 / max(n - 2, 1), 2)}), flush=True)
    ref = ROOT / "oracle" / "_ref" / "treeator"
    if args.ref_taxa and ref.exists():
        n = args.ref_taxa
        tri = matrix(n)
        with tempfile.TemporaryDirectory() as td:
            path = Path(td) / "m.txt"
            write_matrix(path, tri, n)
            t0 = time.perf_counter()
            r = subprocess.run([str(ref), "-n", str(path)], capture_output=True)
            dt = time.perf_counter() - t0
            exe = ROOT / "build" / "treeator_b200"
            t0 = time.perf_counter()
            g = subprocess.run([str(exe), "-n", str(path)], capture_output=True)
            dg = time.perf_counter() - t0
        print(json.dumps({"reference_treeator_n": n, "seconds": round(dt, 3), "treeator_b200_seconds": round(dg, 3),
                          "identical_output": r.stdout == g.stdout and len(r.stdout) > 0, "cores": 1}), flush=True)
    capi.shutdown()


if __name__ == "__main__":
    main()
