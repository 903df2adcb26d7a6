#!/usr/bin/env python3
"""Generate tests/golden/nj/: output of the UNMODIFIED reference `treeator -n` (oracle/_ref/treeator, built by
oracle/Makefile from /root/reference/src) for distance matrices in pairalign's -m text format.
Runs only in the authoring container."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "treeator"
OUT = ROOT / "tests" / "golden" / "nj"
CLI = ROOT / "tests" / "golden" / "cli"


def write_matrix(path, names, d, labels=True, last_name=True, fmt="%g"):
    """pairalign's -n -m layout: row r is 'name ' + r spaces + values each followed by a space; last name alone."""
    n = len(names)
    with open(path, "w") as fh:
        for a in range(n - 1):
            if a:
                fh.write("\n")
            if labels:
                fh.write(names[a] + " ")
            fh.write(" " * a)
            for b in range(a + 1, n):
                fh.write((fmt % d[a, b]) + " ")
        if last_name:
            fh.write("\n" + (names[-1] if labels else "") + "\n")
        else:
            fh.write("\n")


def main():
    if not REF.exists():
        sys.exit("oracle/_ref/treeator missing: run `make -C oracle` where /root/reference exists")
    OUT.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(7)
    cases = []
    # matrices printed by the reference pairalign itself
    for tag in ("example_A_jnm", "mixed_long_flags", "pure_j_n_m", "mixed_s_m_n"):
        (OUT / f"{tag}.matrix").write_bytes((CLI / f"{tag}.out").read_bytes())
        cases.append((tag, ["-n"]))
    # synthetic: tree-like, random, and integer-valued with many exact ties
    for n, kind in ((3, "rand"), (4, "ties"), (5, "rand"), (17, "ties"), (60, "tree"), (250, "rand"), (250, "ties"), (400, "tree")):
        if kind == "tree":
            pts = rng.random((n, 6))
            d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)) + 0.05 * rng.random((n, n))
        elif kind == "rand":
            d = rng.random((n, n)) * 2
        else:
            d = rng.integers(1, 5, size=(n, n)).astype(float) / 4
        d = np.triu(d, 1); d = d + d.T
        names = [f"t{k:04d}" for k in range(n)]
        tag = f"synth_{kind}_{n}"
        write_matrix(OUT / f"{tag}.matrix", names, d)
        cases.append((tag, ["-n"]))
    # option variants on one matrix
    n = 12
    d = rng.random((n, n)); d = np.triu(d, 1); d = d + d.T
    names = [f"x{k}" for k in range(n)]
    write_matrix(OUT / "opt_nolabel.matrix", names, d, labels=False, last_name=False)
    cases.append(("opt_nolabel", ["-n", "-L"]))
    write_matrix(OUT / "opt_nobr.matrix", names, d)
    cases.append(("opt_nobr", ["-n", "-0"]))
    write_matrix(OUT / "opt_sci.matrix", names, d * 1e-5, fmt="%.3e")
    cases.append(("opt_sci", ["--neighbour_joining"]))
    manifest = []
    for tag, flags in cases:
        r = subprocess.run([str(REF), *flags, f"{tag}.matrix"], cwd=OUT, capture_output=True)
        (OUT / f"{tag}.newick").write_bytes(r.stdout)
        manifest.append(dict(tag=tag, flags=flags, rc=r.returncode))
        print(tag, r.returncode, len(r.stdout), r.stderr[:80])
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
