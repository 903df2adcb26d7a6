#!/usr/bin/env python3
"""Generate tests/golden/seqpair_vectors.json from the UNMODIFIED reference.

Runs only in the authoring container (needs oracle/_ref/libref_seqpair.so, which
oracle/Makefile compiles in place from /root/reference/src).  The reference ships no
golden vectors of its own, so these are the pinned outputs of
its seqpair class: score (the value pairalign discards), hamming distance, similarity
and JC as exact hex floats, and the aligned strings of get_x()/get_y().

    python oracle/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import oracle_lib  # noqa: E402
from phylommand_b200 import synth  # noqa: E402


def main():
    ref = oracle_lib.load_ref()
    if ref is None:
        sys.exit("oracle/_ref/libref_seqpair.so missing: run `make -C oracle` where /root/reference exists")
    cases = []

    def add(x, y, aligned=False, tag=""):
        r = ref.run(x, y, aligned)
        cases.append(dict(tag=tag, x=x, y=y, aligned=bool(aligned), score=r["score"], hamming=r["hamming"],
                          sim=float(r["sim"]).hex(), jc=float(r["jc"]).hex(), ax=r["x"], ay=r["y"]))

    # known-answer micro vectors; the leading 'N' is the character the reference drops
    micro = [("ACGTACGTTT", "ACGTCGTTT"), ("CGTCGTTT", "ACGTACGTTTG"), ("AAAA", "TTTT"),
             ("ACGTTTTTTTTTTTTTTTTTTTTACGT", "ACGTACGT"), ("ACGTACGTACGTACGTACGTGGGGG", "CCCCCACGTACGTACGTACGTACGT"),
             ("ACGTNACGTRYACGT", "acgtaacgtagacgu"), ("A", "A"), ("A", "T"), ("ACGT", "A"), ("T", "ACGT"),
             ("AC-GT", "ACGT"), ("A-", "-A"), ("NNNN", "ACGT"), ("RYSWKM", "BDHVN."), ("ACGT ACGT\tAC", "ACGTACGTAC")]
    for x, y in micro:
        add("N" + x, "N" + y, tag="micro")
        add("N" + y, "N" + x, tag="micro-swapped")
    for x, y in micro[:6]:
        add("N" + x, "N" + y, aligned=True, tag="micro-aligned")

    rng = np.random.default_rng(20261017)
    # random pure A/C/G/T pairs, ragged lengths, crossing the 256/512-column pass widths
    for lo, hi, n in ((1, 12, 40), (20, 130, 40), (240, 300, 10), (500, 530, 8), (700, 1100, 6)):
        _, seqs = synth.make_random(2 * n, int(rng.integers(1 << 30)), lo, hi)
        for k in range(n):
            add("N" + synth.to_text(seqs[2 * k]), "N" + synth.to_text(seqs[2 * k + 1]), tag=f"pure-{lo}-{hi}")
    # IUPAC codes
    for lo, hi, n in ((1, 40, 30), (100, 300, 12), (510, 600, 4)):
        _, seqs = synth.make_random(2 * n, int(rng.integers(1 << 30)), lo, hi, iupac=0.05)
        for k in range(n):
            add("N" + synth.to_text(seqs[2 * k]), "N" + synth.to_text(seqs[2 * k + 1]), tag=f"iupac-{lo}-{hi}")
    # '-' in unaligned input: cost() = INT_MIN and 32-bit wrap-around in the reference binary
    for lo, hi, n in ((1, 40, 30), (100, 300, 10)):
        _, seqs = synth.make_random(2 * n, int(rng.integers(1 << 30)), lo, hi, iupac=0.02, gaps=0.15)
        for k in range(n):
            add("N" + synth.to_text(seqs[2 * k]), "N" + synth.to_text(seqs[2 * k + 1]), tag=f"gaps-{lo}-{hi}")
        for k in range(min(n, 6)):
            add("N" + synth.to_text(seqs[2 * k]), "N" + synth.to_text(seqs[2 * k + 1]), aligned=True, tag="gaps-aligned")
    out = ROOT / "tests" / "golden" / "seqpair_vectors.json"
    out.write_text(json.dumps(dict(generator="oracle/make_golden.py", reference="RybergGroup/phylommand src/seqpair.{h,cpp} "
                                   "compiled unmodified (oracle/Makefile)", scoring=dict(match=7, mismatch=-5, gap_open=-15, gap_ext=-1),
                                   cases=cases), indent=0))
    print(f"{len(cases)} cases -> {out} ({out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
