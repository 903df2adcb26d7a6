"""The oracle (oracle/pa_oracle.c) pinned against the reference.

* tests/golden/seqpair_vectors.json: outputs of the UNMODIFIED reference seqpair class
  (made by oracle/make_golden.py) -- score, hamming distance, similarity, JC distance and
  the aligned strings, incl. IUPAC codes and '-' (INT_MIN / 32-bit wrap) inputs.
* when oracle/_ref/libref_seqpair.so is present (authoring container and, prebuilt, the GPU
  box) fresh random pairs are also checked against the live reference.
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from phylommand_b200 import synth
from tests import oracle_lib

GOLDEN = json.loads((Path(__file__).parent / "golden" / "seqpair_vectors.json").read_text())


def _check_case(oracle, case):
    x = oracle.encode(case["x"])
    y = oracle.encode(case["y"])
    if case["aligned"]:
        r = oracle.aligned_stats(x, y)
        assert oracle.decode(x) == case["ax"] and oracle.decode(y) == case["ay"]
    else:
        if len(x) == 0 or len(y) == 0:
            return
        r, ax, ay = oracle.align_full(x, y)
        f = oracle.align_forward(x, y)
        assert tuple(r) == tuple(f), "forward-only form differs from the literal restatement"
        assert int(r["score"]) == case["score"]
        assert oracle.decode(ax) == case["ax"]
        assert oracle.decode(ay) == case["ay"]
    assert int(r["dist"]) == case["hamming"]
    sim = oracle.similarity(int(r["dist"]), int(r["len"]))
    jc = oracle.jc(int(r["dist"]), int(r["len"]))
    assert sim.hex() == case["sim"]
    want_jc = float.fromhex(case["jc"])
    assert (math.isnan(jc) and math.isnan(want_jc)) or jc.hex() == case["jc"]


def test_golden_vectors_cover_the_edge_cases():
    tags = {c["tag"].split("-")[0] for c in GOLDEN["cases"]}
    assert {"micro", "pure", "iupac", "gaps"} <= tags
    assert len(GOLDEN["cases"]) >= 200


@pytest.mark.parametrize("idx", range(len(GOLDEN["cases"])))
def test_oracle_matches_reference_golden(oracle, idx):
    _check_case(oracle, GOLDEN["cases"][idx])


def test_known_answers(oracle):
    # SURVEY.md 8c micro vectors (inputs shown after the first-character drop)
    x, y = oracle.encode("NACGTACGTTT"), oracle.encode("NACGTCGTTT")
    r, ax, ay = oracle.align_full(x, y)
    assert (int(r["score"]), int(r["dist"]), int(r["len"])) == (48, 0, 9)
    assert (oracle.decode(ax), oracle.decode(ay)) == ("ACGTACGTTT", "ACGT-CGTTT")
    x, y = oracle.encode("NCGTCGTTT"), oracle.encode("NACGTACGTTTG")
    r, ax, ay = oracle.align_full(x, y)
    assert (int(r["score"]), int(r["dist"]), int(r["len"])) == (41, 0, 5)
    assert (oracle.decode(ax), oracle.decode(ay)) == ("-----CGTCGTTT-", "ACGTA---CGTTTG")
    x, y = oracle.encode("NAAAA"), oracle.encode("NTTTT")
    r, ax, ay = oracle.align_full(x, y)
    assert (int(r["score"]), int(r["len"])) == (-5, 0)
    assert oracle.similarity(int(r["dist"]), int(r["len"])) == 1.0
    assert (oracle.decode(ax), oracle.decode(ay)) == ("----AAAA", "TTTT----")


def test_encode_rules(oracle):
    # first character dropped, white space skipped, unknown skipped, case-insensitive, N -> '.'
    assert oracle.encode("XACGT").tolist() == [1, 4, 2, 8]
    assert oracle.encode("AAC GT\n").tolist() == [1, 4, 2, 8]
    assert oracle.encode("Nacgu?*t").tolist() == [1, 4, 2, 8]
    assert oracle.decode(oracle.encode("NnN.-RYSWKMBDHV")) == "..." + "-RYSWKMBDHV"
    assert oracle.encode("").tolist() == [] and oracle.encode("A").tolist() == []


def test_stats_special_values(oracle):
    assert math.copysign(1.0, oracle.jc(0, 10)) == -1.0 and oracle.jc(0, 10) == 0.0   # "-0"
    assert math.isinf(oracle.jc(3, 4))                                                # p = 0.75
    assert math.isnan(oracle.jc(4, 4))                                                # p > 0.75
    assert oracle.similarity(0, 0) == 1.0
    assert oracle.pdist(1, 3) == 1 - (1.0 - (1 / 3.0))


def test_forward_equals_full_random(oracle):
    rng = np.random.default_rng(7)
    for trial in range(300):
        kind = trial % 3
        _, seqs = synth.make_random(2, int(rng.integers(1 << 30)), 1, 90,
                                    iupac=0.05 if kind else 0.0, gaps=0.15 if kind == 2 else 0.0)
        x = oracle.encode("N" + synth.to_text(seqs[0]))
        y = oracle.encode("N" + synth.to_text(seqs[1]))
        r, _, _ = oracle.align_full(x, y)
        assert tuple(r) == tuple(oracle.align_forward(x, y))


def test_oracle_against_live_reference(oracle):
    ref = oracle_lib.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built here (the golden vectors above are the committed pin)")
    rng = np.random.default_rng(99)
    for trial in range(150):
        kind = trial % 3
        _, seqs = synth.make_random(2, int(rng.integers(1 << 30)), 1, 160,
                                    iupac=0.04 if kind else 0.0, gaps=0.1 if kind == 2 else 0.0)
        tx, ty = "N" + synth.to_text(seqs[0]), "N" + synth.to_text(seqs[1])
        want = ref.run(tx, ty)
        x, y = oracle.encode(tx), oracle.encode(ty)
        r, ax, ay = oracle.align_full(x, y)
        assert int(r["score"]) == want["score"] and int(r["dist"]) == want["hamming"]
        assert oracle.decode(ax) == want["x"] and oracle.decode(ay) == want["y"]
        assert oracle.similarity(int(r["dist"]), int(r["len"])) == want["sim"]


def test_oracle_against_live_reference_structured(oracle):
    """Inputs where ties decide everything: repeats and homopolymers, one sequence inside the other, overlapping ends,
    nothing but ambiguity codes, one to four bases with gap characters (4 000 such pairs were run once: no difference)."""
    ref = oracle_lib.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built here (the golden vectors above are the committed pin)")
    rng = np.random.default_rng(2024)
    alph, amb = "ACGT", "RYSWKMBDHVN"
    rnd = lambda n, letters=alph: "".join(letters[int(k)] for k in rng.integers(0, len(letters), size=n))
    for trial in range(500):
        kind = trial % 5
        if kind == 0:
            unit = rnd(int(rng.integers(1, 4)))
            tx, ty = "N" + unit * int(rng.integers(1, 40)), "N" + unit * int(rng.integers(1, 40))
            if rng.random() < 0.5:
                ty = ty[:-1] + rnd(1)
        elif kind == 1:
            a = rnd(int(rng.integers(20, 150)))
            i = int(rng.integers(0, len(a) - 3)); j = int(rng.integers(i + 2, len(a)))
            tx, ty = "N" + a, "N" + a[i:j]
            if rng.random() < 0.5:
                tx, ty = ty, tx
        elif kind == 2:
            a = rnd(120)
            k1 = int(rng.integers(10, 110))
            tx, ty = "N" + a[:k1 + int(rng.integers(0, 10))], "N" + a[k1 - int(rng.integers(0, 10)):]
        elif kind == 3:
            tx, ty = "N" + rnd(int(rng.integers(1, 60)), amb + alph), "N" + rnd(int(rng.integers(1, 60)), amb + alph)
        else:
            tx, ty = "N" + rnd(int(rng.integers(1, 5))), "n" + rnd(int(rng.integers(1, 5)), "acgt-")
        want = ref.run(tx, ty)
        x, y = oracle.encode(tx), oracle.encode(ty)
        r, ax, ay = oracle.align_full(x, y)
        assert int(r["score"]) == want["score"] and int(r["dist"]) == want["hamming"], (tx, ty)
        assert oracle.decode(ax) == want["x"] and oracle.decode(ay) == want["y"], (tx, ty)
        assert oracle.similarity(int(r["dist"]), int(r["len"])) == want["sim"], (tx, ty)
        assert tuple(r) == tuple(oracle.align_forward(x, y)), (tx, ty)


def test_compact_op_strings_equal_the_literal_ones(oracle):
    """pa_oracle_align_ops_compact (2-bit moves, for 30 kb pairs) against the full-matrix walk: same record, same ops,
    and the ops render to the gapped strings of align_full."""
    rng = np.random.default_rng(11)
    for trial in range(200):
        kind = trial % 3
        _, seqs = synth.make_random(2, int(rng.integers(1 << 30)), 1, 150,
                                    iupac=0.05 if kind else 0.0, gaps=0.12 if kind == 2 else 0.0, related=bool(trial % 2))
        x = oracle.encode("N" + synth.to_text(seqs[0]))
        y = oracle.encode("N" + synth.to_text(seqs[1]))
        if len(x) == 0 or len(y) == 0:
            continue
        r_full, ax, ay = oracle.align_full(x, y)
        r1, ops1 = oracle.align_ops(x, y)
        r2, ops2 = oracle.align_ops(x, y, compact=True)
        assert tuple(r1) == tuple(r_full) == tuple(r2)
        assert ops1.tolist() == ops2.tolist()
        i = j = 0
        for k, o in enumerate(ops1):
            assert (int(ax[k]) if o != 2 else 0) == (int(x[i]) if o != 2 else 0)
            assert (int(ay[k]) if o != 1 else 0) == (int(y[j]) if o != 1 else 0)
            i += o != 2
            j += o != 1
        assert i == len(x) and j == len(y) and len(ax) == len(ops1)


def test_all_pairs_driver_order(oracle):
    _, seqs = synth.make_random(7, 3, 5, 40)
    enc = [synth.to_masks(s) for s in seqs]
    offsets = np.zeros(8, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(e) for e in enc])
    masks = np.concatenate(enc)
    out = oracle.all_pairs(masks, offsets, threads=3)
    k = 0
    for a in range(7):
        for b in range(a + 1, 7):
            assert tuple(out[k]) == tuple(oracle.align_forward(enc[a], enc[b]))
            k += 1
    part = oracle.all_pairs(masks, offsets, first=5, last=17, threads=2)
    assert [tuple(r) for r in part] == [tuple(r) for r in out[5:17]]
