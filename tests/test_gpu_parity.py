"""Parity of the CUDA path (through the C-ABI) with the oracle.  Bit-exact for every
integer output; floating-point figures are derived on the host from the integers with the
reference's expressions, so they are compared bit-for-bit too (tolerance 0, which is inside
the 1e-12 relative bound BASELINE.json states for JC distances)."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from phylommand_b200 import synth

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((Path(__file__).parent / "golden" / "seqpair_vectors.json").read_text())


def _same(got, want):
    bad = [k for k in range(len(want)) if tuple(got[k]) != tuple(want[k])]
    assert not bad, f"{len(bad)} of {len(want)} pairs differ; first {bad[0]}: got {got[bad[0]]}, want {want[bad[0]]}"


def _oracle_all(oracle, enc, threads=8, **kw):
    masks, offsets = np.concatenate(enc), np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.uint64)
    return oracle.all_pairs(masks, offsets, threads=threads, **kw)


def test_golden_vectors_on_gpu(gpu, oracle):
    """Every reference golden vector (unaligned and -A), one upload, explicit pair list."""
    seqs, ia, ib, cases = [], [], [], []
    for c in GOLDEN["cases"]:
        x, y = gpu.encode(c["x"]), gpu.encode(c["y"])
        if len(x) == 0 or len(y) == 0:
            continue
        ia.append(len(seqs)); seqs.append(x)
        ib.append(len(seqs)); seqs.append(y)
        cases.append(c)
    gpu.upload(seqs)
    ia, ib = np.array(ia), np.array(ib)
    dp = gpu.align_pairs(ia, ib)
    al = gpu.align_pairs(ia, ib, aligned=1)
    for k, c in enumerate(cases):
        r = al[k] if c["aligned"] else dp[k]
        if not c["aligned"]:
            assert int(r["score"]) == c["score"], c["tag"]
        assert int(r["dist"]) == c["hamming"], c["tag"]
        assert gpu.similarity(int(r["dist"]), int(r["len"])).hex() == c["sim"], c["tag"]
        jc, want = gpu.jc_distance(int(r["dist"]), int(r["len"])), float.fromhex(c["jc"])
        assert (math.isnan(jc) and math.isnan(want)) or jc.hex() == c["jc"], c["tag"]


def test_golden_alignments_on_gpu(gpu):
    """pairalign -a: aligned strings equal the reference's get_x()/get_y()."""
    n = 0
    for c in GOLDEN["cases"]:
        if c["aligned"]:
            continue
        x, y = gpu.encode(c["x"]), gpu.encode(c["y"])
        if len(x) == 0 or len(y) == 0 or len(x) * len(y) > 400 * 400:
            continue
        gpu.upload([x, y])
        ax, ay, res = gpu.align_pair_traceback(0, 1, len(x) + len(y))
        assert gpu.decode(ax) == c["ax"] and gpu.decode(ay) == c["ay"], c["tag"]
        assert int(res["score"]) == c["score"]
        n += 1
    assert n > 100


@pytest.mark.parametrize("lo,hi,n,seed", [(1, 40, 60, 1), (30, 200, 40, 2), (480, 560, 14, 3), (1000, 1100, 8, 4)])
def test_all_pairs_pure_acgt(gpu, oracle, lo, hi, n, seed):
    """2-bit register-wavefront kernel: ragged lengths, 1 to 3 passes of 512 columns."""
    _, seqs = synth.make_random(n, seed, lo, hi)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_general_ms"] == 0.0   # s16x2 kernel only
    _same(got, _oracle_all(oracle, enc))
    # the same pairs as an explicit list run on the 32-bit one-pair-per-warp kernel
    ab = np.array([gpu.pair_from_index(k) for k in range(len(got))])
    got32 = gpu.align_pairs(ab[:, 0], ab[:, 1])
    t = gpu.timing()
    assert t["dp_duo_ms"] == 0.0 and t["dp_fast_ms"] > 0
    assert got32.tobytes() == got.tobytes()


def test_every_strip_width_of_the_s16x2_kernel(gpu, oracle):
    """The s16x2 kernel picks its strip width (8, 10, 11, 12 or 13 columns per lane) per work item from the longer
    column sequence: lengths on both sides of every switch point, every width, one to five passes."""
    rng = np.random.default_rng(11)
    lengths = [1, 200, 256, 257, 320, 321, 352, 353, 384, 385, 416, 417, 512, 513, 640, 641, 704, 705, 768, 769, 832,
               833, 960, 1056, 1248, 1536, 1537, 1664, 1665]
    root = synth.BASES[rng.integers(0, 4, size=max(lengths))]
    enc = []
    for L in lengths:
        s = root[:L].copy()
        k = rng.random(L) < 0.08 * rng.random()
        s[k] = synth.BASES[rng.integers(0, 4, size=int(k.sum()))]
        enc.append(synth.to_masks(s))
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_general_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc))


@pytest.mark.parametrize("iupac,gaps,lo,hi,n,seed", [(0.05, 0.0, 1, 60, 40, 5), (0.03, 0.0, 200, 300, 16, 6),
                                                      (0.02, 0.15, 1, 80, 40, 7), (0.0, 0.2, 250, 330, 12, 8)])
def test_all_pairs_iupac_and_gaps(gpu, oracle, iupac, gaps, lo, hi, n, seed):
    """General kernel: IUPAC sets, and '-' in unaligned input (INT_MIN cost, 32-bit wrap-around)."""
    _, seqs = synth.make_random(n, seed, lo, hi, iupac=iupac, gaps=gaps)
    enc = [gpu.encode("N" + synth.to_text(s)) for s in seqs]
    enc = [e if len(e) else np.array([1], dtype=np.uint8) for e in enc]
    gpu.upload(enc)
    _same(gpu.align_all_pairs(), _oracle_all(oracle, enc))


def test_mixed_pure_and_ambiguous(gpu, oracle):
    """Pure pairs run on the 2-bit (PRMT) form of the s16x2 kernel, pairs with IUPAC codes on its 4-bit-set form in the
    same call; only a gap character sends a pair to the general int32 kernel."""
    _, a = synth.make_random(12, 11, 40, 700)
    _, b = synth.make_random(12, 12, 40, 700, iupac=0.03)
    enc = [gpu.encode("N" + synth.to_text(s)) for pair in zip(a, b) for s in pair]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] == 0.0 and t["kernel_launches"] == 3     # work items + the two s16x2 forms
    _same(got, _oracle_all(oracle, enc))
    _, c = synth.make_random(4, 13, 40, 700, iupac=0.03, gaps=0.02)
    enc += [gpu.encode("N" + synth.to_text(s)) for s in c]
    enc = [e if len(e) else np.array([1], dtype=np.uint8) for e in enc]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] > 0
    _same(got, _oracle_all(oracle, enc))


def _sparse_ambiguous(rng, base, n_ranges, first=False):
    """Upper-case sequence with IUPAC ranges the s16x2 kernel accepts: none at positions 1..15, ranges of different
    codes at least 16 plain bases apart (the encoder drops the leading character, so +1 in text coordinates)."""
    s = base.copy()
    amb = np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)
    if first:
        s[0] = amb[rng.integers(0, len(amb))]
    pos = 16 + int(rng.integers(0, 40))
    for _ in range(n_ranges):
        run = int(rng.choice([1, 1, 1, 2, 3, 17, 40]))
        if pos + run >= len(s):
            break
        s[pos:pos + run] = amb[rng.integers(0, len(amb))]
        pos += run + 16 + int(rng.integers(0, max(1, len(s) // (n_ranges + 1))))
    return s


def test_sparse_ambiguity_codes_stay_on_the_s16x2_kernel(gpu, oracle):
    """Sequences with sparse IUPAC codes on the 4-bit-set form of the s16x2 kernel: runs of N that span several lanes,
    an ambiguous first base, ambiguous rows and columns meeting, every strip position, short partners, and a pair
    above the 16-bit limit (floating window + ambiguity)."""
    rng = np.random.default_rng(4242)
    root = synth.BASES[rng.integers(0, 4, size=1800)]
    enc = []
    for k in range(14):
        L = int(rng.integers(60, 1800))
        base = root[:L].copy()
        mut = rng.random(L) < 0.1 * rng.random()
        base[mut] = synth.BASES[rng.integers(0, 4, size=int(mut.sum()))]
        enc.append(synth.to_masks(_sparse_ambiguous(rng, base, int(rng.integers(0, 6)), first=(k % 4 == 0))))
    enc.append(synth.to_masks(root[:700]))                                   # a pure one among them
    enc.append(synth.to_masks(_sparse_ambiguous(rng, root[:20].copy(), 1)))  # shorter than one strip
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    # two launches of the s16x2 kernel: the plain items, then the items with an ambiguous sequence
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_general_ms"] == 0.0 and t["kernel_launches"] == 3     # work items + the two s16x2 forms
    _same(got, _oracle_all(oracle, enc))
    # long pairs: floating window and ambiguity together
    _, longs = synth.make_long(3, 99, length=5200, spread=0.1, div_lo=0.0, div_hi=0.08)
    enc = [synth.to_masks(_sparse_ambiguous(rng, s, 5, first=(k == 1))) for k, s in enumerate(longs)]
    enc.append(synth.to_masks(_sparse_ambiguous(rng, longs[0][:900].copy(), 3)))
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_general_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc, threads=8))
    # the general kernel (explicit pair list) agrees
    ab = np.array([gpu.pair_from_index(k) for k in range(len(got))])
    assert gpu.align_pairs(ab[:, 0], ab[:, 1]).tobytes() == got.tobytes()


def test_dense_ambiguity_codes_stay_on_the_s16x2_kernel(gpu, oracle):
    """Any density of IUPAC codes runs on the 4-bit-set form of the s16x2 kernel (every code, neighbouring different
    codes, codes in the first positions, all-N and all-ambiguous sequences, 1-base sequences, lengths around the strip
    and pass widths, odd / even row counts); a gap character alone sends a pair to the general int32 kernel."""
    rng = np.random.default_rng(5)
    base = synth.BASES[rng.integers(0, 4, size=400)]
    a = base.copy(); a[100] = ord("R"); a[105] = ord("Y")
    b = base.copy(); b[3] = ord("N")
    c = base.copy(); c[200] = ord("N"); c[300] = ord("N")
    enc = [synth.to_masks(x) for x in (a, b, c, base)]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] == 0.0 and t["dp_fast_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc))
    enc.append(gpu.encode("N" + synth.to_text(base[:150]) + "-" + synth.to_text(base[150:300])))
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] > 0
    _same(got, _oracle_all(oracle, enc))
    # dense random codes, all lengths
    amb = np.frombuffer(b"ACGTRYSWKMBDHVN", dtype=np.uint8)
    enc = [synth.to_masks(amb[rng.integers(0, len(amb), size=int(L))]) for L in
           (1, 2, 3, 11, 12, 13, 24, 25, 383, 384, 385, 386, 700, 767, 768, 769, 1153, 40, 97)]
    enc.append(np.full(300, 15, dtype=np.uint8))                    # all N: every cell a match
    enc.append(synth.to_masks(np.frombuffer(b"RY" * 200, dtype=np.uint8)))   # R/Y alternating: never intersect each other
    _, rel = synth.make_random(6, 77, 200, 900, iupac=0.3)
    enc += [gpu.encode("N" + synth.to_text(s)) for s in rel]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] == 0.0 and t["dp_fast_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc))
    # sub-ranges of the triangle (a half-wanted work item mirrors its other half) and other scoring parameters
    half = gpu.num_pairs() // 2
    assert gpu.align_all_pairs(3, half).tobytes() == got[3:3 + half].tobytes()
    for sc in (dict(match=5, mismatch=-4, gap_open=-10, gap_ext=-2), dict(match=3, mismatch=1, gap_open=-7, gap_ext=-1),
               dict(match=2, mismatch=-9, gap_open=0, gap_ext=0), dict(match=-1, mismatch=-3, gap_open=-4, gap_ext=-2)):
        got = gpu.align_all_pairs(**sc)
        _same(got, _oracle_all(oracle, enc, match=sc["match"], mismatch=sc["mismatch"], go=sc["gap_open"], ge=sc["gap_ext"]))
    # long ambiguous pairs: floating window on the set form
    _, longs = synth.make_long(3, 98, length=5600, spread=0.1, div_lo=0.0, div_hi=0.08)
    enc = []
    for s in longs:
        s = s.copy()
        k = rng.random(len(s)) < 0.05
        s[k] = amb[rng.integers(4, len(amb), size=int(k.sum()))]
        enc.append(synth.to_masks(s))
    enc.append(enc[0][:1000].copy())
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_general_ms"] == 0.0 and t["dp_fast_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc, threads=8))


def test_other_scoring_parameters(gpu, oracle):
    _, seqs = synth.make_random(14, 13, 20, 300, iupac=0.01)
    enc = [gpu.encode("N" + synth.to_text(s)) for s in seqs]
    gpu.upload(enc)
    for match, mismatch, go, ge in ((7, -5, -15, -1), (5, -4, -10, -2), (1, -1, -2, 0), (10, -300, -20, -3), (2, -3, 4, 1)):
        got = gpu.align_all_pairs(match=match, mismatch=mismatch, gap_open=go, gap_ext=ge)
        _same(got, _oracle_all(oracle, enc, match=match, mismatch=mismatch, go=go, ge=ge))


def test_sub_ranges_and_pair_lists(gpu, oracle):
    _, seqs = synth.make_random(25, 21, 10, 120)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    want = _oracle_all(oracle, enc)
    total = gpu.num_pairs()
    assert total == 300
    _same(gpu.align_all_pairs(17, 101), want[17:118])
    _same(gpu.align_all_pairs(total - 1, 1), want[total - 1:])
    assert len(gpu.align_all_pairs(5, 0)) == 0
    idx = np.random.default_rng(0).permutation(total)[:77]
    ab = np.array([gpu.pair_from_index(int(k)) for k in idx])
    _same(gpu.align_pairs(ab[:, 0], ab[:, 1]), want[idx])
    # both orders of the same pair are separate problems (end-cell tie rules are not symmetric)
    rev = gpu.align_pairs(ab[:, 1], ab[:, 0])
    for k in range(len(idx)):
        assert tuple(rev[k]) == tuple(oracle.align_forward(enc[ab[k, 1]], enc[ab[k, 0]]))
    with pytest.raises(gpu.PairalignError):
        gpu.align_all_pairs(total, 1)
    with pytest.raises(gpu.PairalignError):
        gpu.align_pairs([0], [25])


def test_two_device_entries_share_a_call(gpu, oracle):
    """The in-process multi-device path at the C-ABI: every device entry takes a cell-balanced share of a range, of a
    pair list and of a batch of alignments; the records are those of one device.  On a one-GPU box the second entry
    is the same chip (the code path is the same)."""
    import torch
    second = 1 if torch.cuda.device_count() >= 2 else 0
    _, seqs = synth.make_random(40, 23, 50, 900)
    _, amb = synth.make_random(4, 24, 50, 600, iupac=0.02)
    enc = [synth.to_masks(s) for s in seqs] + [gpu.encode("N" + synth.to_text(s)) for s in amb]
    want = _oracle_all(oracle, enc)
    try:
        gpu.init([0, second])
        gpu.upload(enc)
        _same(gpu.align_all_pairs(), want)
        assert gpu.timing()["n_devices"] == 2
        _same(gpu.align_all_pairs(100, 300), want[100:400])
        ia, ib = np.array([0, 5, 41, 7, 3, 20, 11]), np.array([9, 2, 6, 40, 30, 21, 12])
        stats = gpu.align_pairs(ia, ib)
        for k in range(len(ia)):
            assert tuple(stats[k]) == tuple(oracle.align_forward(enc[ia[k]], enc[ib[k]]))
        lens = np.array([len(e) for e in enc])
        ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
        assert res.tobytes() == stats.tobytes()
        for k in range(len(ia)):
            r, wx, wy = oracle.align_full(enc[ia[k]], enc[ib[k]])
            ax, ay = _render(ops[int(off[k]):int(off[k]) + int(n_ops[k])], enc[ia[k]], enc[ib[k]])
            assert ax == wx.tolist() and ay == wy.tolist()
        gpu.upload(enc[:20])                          # a second, smaller set on both
        _same(gpu.align_all_pairs(), _oracle_all(oracle, enc[:20]))
    finally:
        gpu.init()                                    # the session's one-device context for the tests that follow


def test_aligned_mode(gpu, oracle):
    """pairalign -A: position-wise over min(n,m) columns, gaps and IUPAC included."""
    _, seqs = synth.make_random(20, 31, 1, 300, iupac=0.05, gaps=0.2)
    enc = [gpu.encode("N" + synth.to_text(s)) for s in seqs]
    enc = [e if len(e) else np.array([0], dtype=np.uint8) for e in enc]
    gpu.upload(enc)
    got = gpu.align_all_pairs(aligned=1)
    k = 0
    for a in range(20):
        for b in range(a + 1, 20):
            assert tuple(got[k]) == tuple(oracle.aligned_stats(enc[a], enc[b])), (a, b)
            k += 1


def test_traceback_matches_oracle(gpu, oracle):
    _, seqs = synth.make_random(10, 41, 5, 330, iupac=0.02, gaps=0.03)
    enc = [gpu.encode("N" + synth.to_text(s)) for s in seqs]
    gpu.upload(enc)
    for a in range(10):
        for b in (a + 1, (a + 5) % 10):
            if a == b or b >= 10:
                continue
            ax, ay, res = gpu.align_pair_traceback(a, b, len(enc[a]) + len(enc[b]))
            r, wx, wy = oracle.align_full(enc[a], enc[b])
            assert tuple(res) == tuple(r)
            assert ax.tolist() == wx.tolist() and ay.tolist() == wy.tolist()


def test_partition_is_balanced_and_exact(gpu):
    names, seqs, _ = synth.make_its_like(300, 1004)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    total = gpu.num_pairs()
    lens = np.array([len(e) for e in enc], dtype=np.int64)
    cells = int(((lens.sum() ** 2) - (lens ** 2).sum()) // 2)
    assert gpu.count_cells(0, total) == cells
    for parts in (1, 2, 4, 8):
        b = gpu.partition_pairs(0, total, parts)
        assert b[0] == 0 and b[-1] == total and np.all(np.diff(b.astype(np.int64)) >= 0)
        shares = [gpu.count_cells(int(b[p]), int(b[p + 1] - b[p])) for p in range(parts)]
        assert sum(shares) == cells
        assert max(shares) - min(shares) <= 2 * int(lens.max()) ** 2


def test_size_independent_properties_at_scale(gpu, oracle):
    """Config-2-sized sequences (1.5 kb): self-consistency on the full set, oracle on a sample."""
    names, seqs = synth.make_16s_like(96, 1002)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc + [enc[0]])          # a duplicate of sequence 0 at the end
    got = gpu.align_all_pairs()
    n = len(enc) + 1
    lens = np.array([len(e) for e in enc] + [len(enc[0])])
    k = 0
    for a in range(n):
        for b in range(a + 1, n):
            r = got[k]
            assert r["dist"] <= r["len"] <= min(lens[a], lens[b])
            assert r["score"] <= 7 * min(lens[a], lens[b])
            assert (r["end_i"] == lens[a] - 1) or (r["end_j"] == lens[b] - 1)
            if a == 0 and b == n - 1:      # identical sequences: full-length diagonal
                assert (r["score"], r["dist"], r["len"]) == (7 * lens[0], 0, lens[0])
            k += 1
    sample = np.random.default_rng(1).choice(len(got), size=120, replace=False)
    for q in sample:
        a, b = gpu.pair_from_index(int(q))
        ea = enc[a] if a < len(enc) else enc[0]
        eb = enc[b] if b < len(enc) else enc[0]
        assert tuple(got[q]) == tuple(oracle.align_forward(ea, eb)), (a, b)


def test_long_pair_many_passes(gpu, oracle):
    """A 9 kb x 7 kb pair: 14-18 passes, sequences read from global memory instead of the staging buffer."""
    _, seqs = synth.make_long(2, 1005, length=8000, spread=0.12)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    assert tuple(got[0]) == tuple(oracle.align_forward(enc[0], enc[1]))


def test_s16x2_kernel_near_its_int16_limit(gpu, oracle):
    """4 kb sequences: scores reach 28 000 (identical pair); stored with the kernel's negative bias they come within
    a few dozen of 0 from below, the other end of the range stays inside int16 (max_len16 = 4059 for pairalign's
    scoring).  A longer sequence switches its work items to the floating-window variant of the same kernel."""
    _, seqs = synth.make_long(3, 77, length=3970, spread=0.02, div_lo=0.0, div_hi=0.05)
    enc = [synth.to_masks(s) for s in seqs]
    enc.append(enc[0].copy())                       # identical to sequence 0: the highest possible score
    rng = np.random.default_rng(79)
    enc.append(synth.to_masks(synth.BASES[rng.integers(0, 4, size=4059)]))   # exactly the limit, unrelated to the others
    enc.append(np.full(4059, 1, dtype=np.uint8))    # poly-A at the limit: identical pair below scores 7 * 4059
    enc.append(np.full(4059, 1, dtype=np.uint8))
    enc.append(np.full(4058, 8, dtype=np.uint8))    # poly-T: every column a mismatch against poly-A
    _, longer = synth.make_long(1, 78, length=4700, spread=0.0)
    enc.append(synth.to_masks(longer[0]))           # above max_len16: floating-window variant
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_cta_ms"] == 0.0
    _same(got, _oracle_all(oracle, enc, threads=12))
    assert int(got[2]["score"]) == 7 * len(enc[0])  # pair (0, 3)
    scores = {(a, b): int(got[k]["score"]) for k, (a, b) in enumerate((a, b) for a in range(len(enc)) for b in range(a + 1, len(enc)))}
    assert scores[(5, 6)] == 7 * 4059


def test_floating_window_s16x2_long_pairs(gpu, oracle):
    """Pairs of 5-8 kb on the s16x2 kernel: true scores far outside int16 (an identical 8 kb pair scores 56 000, an
    unrelated one goes negative), every lane re-bases its 16-bit window many times, edge rows carry their offsets
    from pass to pass; mixed with short partners, odd lengths, and a sequence against its own prefix."""
    rng = np.random.default_rng(808)
    _, related = synth.make_long(3, 808, length=7000, spread=0.12, div_lo=0.0, div_hi=0.1)
    enc = [synth.to_masks(s) for s in related]
    enc.append(enc[0].copy())                                            # identical: 7 * len
    enc.append(synth.to_masks(synth.BASES[rng.integers(0, 4, size=8191)]))   # unrelated, odd length
    enc.append(synth.to_masks(synth.BASES[rng.integers(0, 4, size=5001)]))
    enc.append(enc[1][:4000].copy())                                     # prefix of a long one (short partner)
    enc.append(enc[2][1500:1537].copy())                                 # a 37-base piece
    enc.append(enc[0][::-1].copy())                                      # reversed: no long diagonal at all
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_cta_ms"] == 0.0 and t["kernel_launches"] == 2     # work items + DP
    _same(got, _oracle_all(oracle, enc, threads=12))
    assert int(got[2]["score"]) == 7 * len(enc[0]) > 32767               # pair (0, 3)
    # the int32 one-pair-per-warp kernel (explicit pair list) agrees
    ab = np.array([gpu.pair_from_index(k) for k in range(len(got))])
    assert gpu.align_pairs(ab[:, 0], ab[:, 1]).tobytes() == got.tobytes()
    # other scoring parameters on the same path (wider gaps between neighbouring states, lower 16-bit limit)
    for sc in (dict(match=5, mismatch=-4, gap_open=-10, gap_ext=-2), dict(match=20, mismatch=-25, gap_open=-100, gap_ext=-7)):
        gpu.upload(enc[:6])
        got = gpu.align_all_pairs(**sc)
        t = gpu.timing()
        assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] == 0.0 and t["dp_cta_ms"] == 0.0
        _same(got, _oracle_all(oracle, enc[:6], threads=12, match=sc["match"], mismatch=sc["mismatch"], go=sc["gap_open"],
                               ge=sc["gap_ext"]))


def test_cta_per_pair_kernel_long_pairs(gpu, oracle):
    """Pairs longer than 8192 take a whole CTA: 8 warps on consecutive 512-column blocks of one pair,
    edges through shared-memory rings (and global memory for the wrap-around).  18-24 blocks = 3 rounds."""
    _, seqs = synth.make_long(3, 501, length=10500, spread=0.12)
    enc = [synth.to_masks(s) for s in seqs]
    enc.append(enc[1][:300].copy())                  # long x against a one-block y (and the other way round)
    enc.append(enc[0][:8193].copy())                 # just above the threshold, odd length
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_cta_ms"] > 0
    _same(got, _oracle_all(oracle, enc, threads=10))
    # explicit list, both orientations
    ia, ib = np.array([3, 0, 4, 2]), np.array([0, 3, 2, 4])
    rev = gpu.align_pairs(ia, ib)
    for k in range(len(ia)):
        assert tuple(rev[k]) == tuple(oracle.align_forward(enc[ia[k]], enc[ib[k]])), (ia[k], ib[k])


def test_few_long_pairs_take_the_cta_per_item_route(gpu, oracle, monkeypatch):
    """A triangle range of long A/C/G/T pairs too small to fill the warp-per-item statistics kernel (config 5 cut over
    several GPUs) runs on the move-storing CTA-per-item kernel, the walk counting the statistics: records equal the
    oracle's and the warp kernel's (floating window) on the same pairs; sub-ranges and device-resident output too."""
    _, seqs = synth.make_long(5, 901, length=9000, spread=0.08)
    enc = [synth.to_masks(s) for s in seqs]
    enc.append(enc[0][:8300].copy())
    want = _oracle_all(oracle, enc, threads=12)
    gpu.upload(enc)                      # 15 pairs: far too few for a warp each
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_cta_ms"] > 0 and t["walk_ms"] > 0 and t["dp_duo_ms"] == 0.0
    _same(got, want)
    _same(gpu.align_all_pairs(2, 9), want[2:11])
    monkeypatch.setenv("PAIRALIGN_NO_CTA", "1")          # the warp-per-item kernel on the same pairs
    gpu.init()
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["walk_ms"] == 0.0 and t["dp_cta_ms"] == 0.0
    _same(got, want)
    monkeypatch.delenv("PAIRALIGN_NO_CTA")
    gpu.init()


def test_all_four_dp_kernels_in_one_call(gpu, oracle):
    """s16x2 (short), int32 warp (4.6-8 kb), CTA (> 8 kb) and general (IUPAC) pairs from one upload."""
    _, short = synth.make_random(6, 601, 200, 900)
    _, mid = synth.make_long(2, 602, length=6000, spread=0.1)
    _, long_ = synth.make_long(2, 603, length=9000, spread=0.05)
    _, amb = synth.make_random(3, 604, 300, 700, iupac=0.06)     # dense codes: not for the s16x2 kernel
    enc = [synth.to_masks(s) for s in short + mid + long_] + [gpu.encode("N" + synth.to_text(s)) for s in amb]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_fast_ms"] > 0 and t["dp_cta_ms"] > 0 and t["dp_general_ms"] > 0
    assert t["kernel_launches"] in (5, 6)     # work items, four DP kernels, and the set form of the s16x2 kernel
    _same(got, _oracle_all(oracle, enc, threads=10))


def _render(ops, x, y):
    ax, ay, i, j = [], [], 0, 0
    for o in ops:
        ax.append(x[i] if o != 2 else 0); i += (o != 2)
        ay.append(y[j] if o != 1 else 0); j += (o != 1)
    assert i == len(x) and j == len(y)
    return ax, ay


def test_batched_alignments(gpu, oracle):
    """pairalign -a in batches: op strings for A/C/G/T pairs (int32 kernel with 2-bit moves), IUPAC / gap pairs
    (general kernel) and a multi-block pair, against the reference-literal traceback of the oracle."""
    _, pure = synth.make_random(10, 701, 5, 700)
    _, amb = synth.make_random(6, 702, 5, 400, iupac=0.03, gaps=0.02)
    _, longer = synth.make_long(2, 703, length=2600, spread=0.1)
    enc = [synth.to_masks(s) for s in pure] + [gpu.encode("N" + synth.to_text(s)) for s in amb] + [synth.to_masks(s) for s in longer]
    enc = [e if len(e) else np.array([1], dtype=np.uint8) for e in enc]
    gpu.upload(enc)
    n = len(enc)
    ia, ib = np.array([a for a in range(n) for b in range(n) if a != b]), np.array([b for a in range(n) for b in range(n) if a != b])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    stats = gpu.align_pairs(ia, ib)
    assert res.tobytes() == stats.tobytes()
    for k in range(len(ia)):
        r, wx, wy = oracle.align_full(enc[ia[k]], enc[ib[k]])
        assert tuple(res[k]) == tuple(r), (ia[k], ib[k])
        ax, ay = _render(ops[int(off[k]):int(off[k]) + int(n_ops[k])], enc[ia[k]], enc[ib[k]])
        assert ax == wx.tolist() and ay == wy.tolist(), (ia[k], ib[k])


def test_batched_alignments_with_iupac_codes_on_the_set_form(gpu, oracle):
    """pairalign -a for sequences with IUPAC codes (no gap character): the move-storing s16x2 kernels in their 4-bit-set
    form -- warp per item for short pairs, floating window above the 16-bit limit, CTA per item above 8192 -- against the
    oracle's walk; plain and ambiguous entries mixed in one list, single-entry items, both orientations."""
    rng = np.random.default_rng(77)
    amb = np.frombuffer(b"ACGTRYSWKMBDHVN", dtype=np.uint8)
    _, short = synth.make_random(8, 705, 3, 800, iupac=0.08)
    enc = [gpu.encode("N" + synth.to_text(s)) for s in short]
    enc = [e if len(e) else np.array([15], dtype=np.uint8) for e in enc]
    enc.append(synth.to_masks(amb[rng.integers(0, len(amb), size=513)]))         # dense codes, one column past a block
    _, plain = synth.make_random(2, 706, 100, 600)
    enc += [synth.to_masks(s) for s in plain]
    _, longs = synth.make_long(3, 707, length=5200, spread=0.1, div_lo=0.0, div_hi=0.08)
    for s_ in longs[:2]:
        s_ = s_.copy(); k = rng.random(len(s_)) < 0.03; s_[k] = amb[rng.integers(4, len(amb), size=int(k.sum()))]
        enc.append(synth.to_masks(s_))
    _, very = synth.make_long(3, 708, length=8800, spread=0.05)
    for s_ in very:
        s_ = s_.copy(); k = rng.random(len(s_)) < 0.02; s_[k] = amb[rng.integers(4, len(amb), size=int(k.sum()))]
        enc.append(synth.to_masks(s_))
    gpu.upload(enc)
    n = len(enc)
    short_n = 11
    pairs = [(a, b) for a in range(short_n) for b in range(short_n) if a != b]
    pairs += [(11, 12), (12, 11), (11, 3), (13, 14), (13, 15), (14, 15), (15, 13), (13, 0)]
    ia, ib = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_cta_ms"] > 0 and t["dp_general_ms"] == 0.0 and t["walk_ms"] > 0
    stats = gpu.align_pairs(ia, ib)
    assert res.tobytes() == stats.tobytes()
    for k in range(len(ia)):
        r, want = oracle.align_ops(enc[ia[k]], enc[ib[k]], compact=max(lens[ia[k]], lens[ib[k]]) > 3000)
        assert tuple(res[k]) == tuple(r), (ia[k], ib[k])
        assert ops[int(off[k]):int(off[k]) + int(n_ops[k])].tobytes() == want.tobytes(), (ia[k], ib[k])


def test_batched_alignments_long_pairs_take_a_cta(gpu, oracle):
    """pairalign -a for pairs longer than 8192: pa_cta32_kernel<16, true> (eight warps per pair, moves stored by
    every block) next to the one-pair-per-warp kernel for the shorter pairs of the same batch, walked back by
    the warp-per-pair walk kernel; against the reference-literal traceback of the oracle."""
    _, seqs = synth.make_long(3, 801, length=9400, spread=0.1)
    enc = [synth.to_masks(s) for s in seqs]
    enc.append(enc[1][:300].copy())                  # one-block y against a long x, and the other way round
    enc.append(enc[2][:8193].copy())                 # just above the threshold, odd length
    _, short = synth.make_random(2, 802, 200, 900)
    enc += [synth.to_masks(s) for s in short]        # stay on the warp kernel
    gpu.upload(enc)
    ia = np.array([0, 1, 3, 0, 4, 5, 2, 6])
    ib = np.array([1, 0, 0, 3, 2, 6, 4, 5])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    t = gpu.timing()
    assert t["dp_cta_ms"] > 0 and t["walk_ms"] > 0 and t["kernel_launches"] == 3
    stats = gpu.align_pairs(ia, ib)
    assert res.tobytes() == stats.tobytes()
    for k in range(len(ia)):
        r, wx, wy = oracle.align_full(enc[ia[k]], enc[ib[k]])
        assert tuple(res[k]) == tuple(r), (ia[k], ib[k])
        ax, ay = _render(ops[int(off[k]):int(off[k]) + int(n_ops[k])], enc[ia[k]], enc[ib[k]])
        assert ax == wx.tolist() and ay == wy.tolist(), (ia[k], ib[k])


def test_alignments_at_config5_size_are_consistent(gpu):
    """BASELINE.json config 5 sizes (30 kb pairs), where the oracle's full matrices do not fit: the op strings must
    consume both sequences exactly, and recounting compared / differing columns along them must give the (dist, len)
    of the record -- which in turn must equal what the statistics-only s16x2 floating-window kernel carried forward."""
    _, seqs = synth.make_long(6, 1005, length=30000, spread=0.05)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    n = len(enc)
    ia = np.array([a for a in range(n) for b in range(a + 1, n)])
    ib = np.array([b for a in range(n) for b in range(a + 1, n)])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    assert gpu.timing()["dp_cta_ms"] > 0
    stats = gpu.align_all_pairs()
    assert res.tobytes() == stats.tobytes()
    for k in range(len(ia)):
        o = ops[int(off[k]):int(off[k]) + int(n_ops[k])]
        x, y = enc[ia[k]], enc[ib[k]]
        assert int((o != 2).sum()) == len(x) and int((o != 1).sum()) == len(y)
        xi = np.cumsum(o != 2) - 1
        yj = np.cumsum(o != 1) - 1
        both = o == 0
        assert int(both.sum()) == int(res[k]["len"])
        assert int(((x[xi[both]] & y[yj[both]]) == 0).sum()) == int(res[k]["dist"])
        # the walk starts at the end cell: no compared column lies beyond it
        last = int(np.flatnonzero(both)[-1])
        assert xi[last] <= int(res[k]["end_i"]) and yj[last] <= int(res[k]["end_j"])
