"""GPU checks at the sizes and on the inputs where parity is hardest (all ran green on a B200 in round 2, first session):

  * the differential fuzz of tests/test_host_fuzz_vs_reference.py through the real command line (CUDA module)
    against the reference binary that travels in oracle/_ref/;
  * the pair-FASTA round trip (`-a -n` read back with `--format pairfst`) the same way;
  * CUDA vs oracle on inputs where ties decide everything (repeats, containment, overlapping ends, all-ambiguous);
  * the op strings of two 30 kb pairs against the oracle's 2-bit-move walk (exact parity of -a at config 5 size);
  * one 30 kb x 30 kb and one 30 kb x 400 bp pair, statistics records against the oracle's forward form;
  * the command line against the reference's own output on the first 64 sequences of config 3 and the first 128 of
    config 4 (tests/golden/prefix/, made by oracle/make_prefix_golden.py), SURVEY.md section 8(d)'s prefix rule."""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from phylommand_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "pairalign"


@pytest.fixture(scope="module")
def cli():
    from phylommand_b200 import build
    build.build_library()
    return build.build_cli()


@pytest.mark.parametrize("seed", range(6))
def test_command_line_fuzz_against_the_reference(cli, tmp_path, seed):
    if not REF.exists():
        pytest.skip("oracle/_ref/pairalign did not travel")
    from tests import test_host_fuzz_vs_reference as F
    env = dict(os.environ, PAIRALIGN_DEVICES="0")
    for text, modes in ((F.make_case(1000 + seed, False), F.MODES[:6]), (F.make_case(2000 + seed, True), F.GROUP_MODES[:4]),
                        (F.make_odd_case(4000 + seed), F.ODD_MODES[:5])):
        (tmp_path / "in.fst").write_bytes(text.encode())
        for flags in modes:
            outs = []
            for binary in (REF, cli):
                r = subprocess.run([str(binary), *flags, "in.fst"], cwd=tmp_path, capture_output=True, timeout=300, env=env)
                g = tmp_path / "in.fst.alignment_groups"
                outs.append((r.returncode, r.stdout, g.read_bytes() if g.exists() else None))
                if g.exists():
                    g.unlink()
            assert outs[0] == outs[1], (flags, seed)


@pytest.mark.parametrize("seed", range(3))
def test_pairfasta_round_trip_through_the_cuda_module(cli, tmp_path, seed):
    """The miniptera.pl pipeline on the GPU: `-a -n` output read back with `--format pairfst`, every
    step against the reference binary doing the same."""
    if not REF.exists():
        pytest.skip("oracle/_ref/pairalign did not travel")
    from tests import test_host_fuzz_vs_reference as F
    env = dict(os.environ, PAIRALIGN_DEVICES="0")

    def both(flags, name):
        outs = []
        for binary in (REF, cli):
            r = subprocess.run([str(binary), *flags, name], cwd=tmp_path, capture_output=True, timeout=300, env=env)
            g = tmp_path / (name + ".alignment_groups")
            outs.append((r.returncode, r.stdout, g.read_bytes() if g.exists() else None))
            if g.exists():
                g.unlink()
        return outs

    (tmp_path / "in.fst").write_bytes(F.make_case(3000 + seed, group=False).encode())
    ref, ours = both(["-a", "-n"], "in.fst")
    assert ours == ref
    (tmp_path / "pairs.pairfst").write_bytes(ours[1])
    for flags in (["--format", "pairfst", "-A", "-j", "-n"], ["--format", "pairfst", "-d", "-n"], ["--format", "pairfst", "-p", "-m", "-n"],
                  ["--format", "pairfst", "-a", "-n"], ["--format", "pairfst", "-g", "alignment_groups"]):
        r, o = both(flags, "pairs.pairfst")
        assert o == r, (flags, seed)


def test_structured_inputs_on_every_kernel(gpu, oracle):
    rng = np.random.default_rng(2024)
    alph, amb = "ACGT", "RYSWKMBDHVN"
    rnd = lambda n, letters=alph: "".join(letters[int(k)] for k in rng.integers(0, len(letters), size=n))
    texts = []
    for trial in range(40):
        kind = trial % 4
        if kind == 0:
            unit = rnd(int(rng.integers(1, 4)))
            texts += [unit * int(rng.integers(1, 200)), unit * int(rng.integers(1, 200))]
        elif kind == 1:
            a = rnd(int(rng.integers(20, 900)))
            i = int(rng.integers(0, len(a) - 3)); j = int(rng.integers(i + 2, len(a)))
            texts += [a, a[i:j]]
        elif kind == 2:
            a = rnd(700)
            k1 = int(rng.integers(10, 690))
            texts += [a[:k1 + int(rng.integers(0, 10))], a[k1 - int(rng.integers(0, 10)):]]
        else:
            texts += [rnd(int(rng.integers(1, 300)), amb + alph), rnd(int(rng.integers(1, 300)), amb + alph)]
    enc = [gpu.encode("N" + t) for t in texts]
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    masks, offsets = gpu.pack(enc)
    want = oracle.all_pairs(masks, offsets, threads=8)
    assert got.tobytes() == want.tobytes()


def test_alignments_of_30kb_pairs_equal_the_oracle(gpu, oracle):
    """BASELINE.json config 5 size, exactly: the op strings of two 30 kb pairs (CTA-per-pair kernel with stored moves,
    warp-per-pair walk) against the oracle's walk over its own 2-bit moves (pa_oracle_align_ops_compact, which
    tests/test_oracle.py ties to the reference-literal full-matrix walk on small pairs).  About 20 s of CPU."""
    _, seqs = synth.make_long(3, 1006, length=30000, spread=0.05)
    enc = [synth.to_masks(s) for s in seqs]
    gpu.upload(enc)
    ia, ib = np.array([0, 2]), np.array([1, 0])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    for k in range(len(ia)):
        r, want = oracle.align_ops(enc[ia[k]], enc[ib[k]], compact=True)
        assert tuple(res[k]) == tuple(r), (ia[k], ib[k])
        got = ops[int(off[k]):int(off[k]) + int(n_ops[k])]
        assert got.tobytes() == want.tobytes(), (ia[k], ib[k])


def test_30kb_pairs_statistics_equal_the_oracle(gpu, oracle):
    """BASELINE.json config 5 size against the oracle (O(n + m) memory forward form): a 30 kb x 30 kb pair and a
    30 kb x 400 bp pair, in both orientations, on every route a long pair can take -- the floating-window s16x2
    kernel (triangle range), the int32 warp kernel / CTA-per-pair kernel (explicit list)."""
    _, seqs = synth.make_long(2, 1005, length=30000, spread=0.05)
    _, short = synth.make_random(1, 1007, 400, 400)
    enc = [synth.to_masks(s) for s in seqs + short]
    gpu.upload(enc)
    from concurrent.futures import ThreadPoolExecutor
    keys = ((0, 1), (0, 2), (1, 2), (1, 0), (2, 0))
    with ThreadPoolExecutor(len(keys)) as ex:       # the oracle is plain C behind ctypes: the calls run side by side
        want = dict(zip(keys, ex.map(lambda ab: tuple(oracle.align_forward(enc[ab[0]], enc[ab[1]])), keys)))
    got = gpu.align_all_pairs()
    assert [tuple(r) for r in got] == [want[(0, 1)], want[(0, 2)], want[(1, 2)]]
    ia, ib = np.array([0, 1, 2, 0]), np.array([1, 0, 0, 2])
    lst = gpu.align_pairs(ia, ib)
    assert [tuple(r) for r in lst] == [want[(0, 1)], want[(1, 0)], want[(2, 0)], want[(0, 2)]]


PREFIX = ROOT / "tests" / "golden" / "prefix"


@pytest.mark.parametrize("tag,flags,groups", [("c3_64", ["-j", "-n", "-m"], False),
                                              ("c4_128", ["--group", "both:cut-off=0.97"], True)])
def test_command_line_on_config_prefixes_matches_the_reference(cli, tmp_path, tag, flags, groups):
    """First 64 sequences of config 3 / first 128 of config 4 (with taxon strings): stdout and the .alignment_groups
    file byte for byte what the unmodified reference printed (about 10 CPU-minutes each there)."""
    shutil.copy(PREFIX / f"{tag}.fst", tmp_path / f"{tag}.fst")
    r = subprocess.run([str(cli), *flags, f"{tag}.fst"], cwd=tmp_path, capture_output=True, timeout=600,
                       env=dict(os.environ, PAIRALIGN_DEVICES="0"))
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-1000:]
    assert r.stdout == (PREFIX / f"{tag}.out").read_bytes()
    if groups:
        assert (tmp_path / f"{tag}.fst.alignment_groups").read_bytes() == (PREFIX / f"{tag}.alignment_groups").read_bytes()


def test_floating_window_on_structured_long_pairs(gpu, oracle, monkeypatch):
    """The 16-bit floating window rests on an analytic bound (what a lane holds lies within a few hundred of its right
    edge); inputs built to stress it -- a 2.5 kb deletion, a 2 kb poly-T insertion, poly-A, a 4-base repeat, the reverse --
    at 9-10 kb, on both routes a long pair can take for its statistics (CTA per item with stored moves + walk; warp per
    item carrying counters) and, for three of the pairs, as op strings."""
    rng = np.random.default_rng(31337)
    a = synth.BASES[rng.integers(0, 4, size=9500)]
    seqs = [a, np.concatenate([a[:3500], a[6000:]]), np.concatenate([a[:4000], np.full(2000, ord("T"), dtype=np.uint8), a[4000:8500]]),
            np.full(9000, ord("A"), dtype=np.uint8), np.frombuffer(b"ACGT" * 2400, dtype=np.uint8), a[::-1].copy()]
    seqs[1] = np.concatenate([seqs[1], synth.BASES[rng.integers(0, 4, size=2000)]])        # keep every sequence above 8192
    enc = [synth.to_masks(x) for x in seqs]
    assert min(len(e) for e in enc) > 8192
    masks, offsets = gpu.pack(enc)
    want = oracle.all_pairs(masks, offsets, threads=15)
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_cta_ms"] > 0 and t["walk_ms"] > 0
    assert got.tobytes() == want.tobytes()
    monkeypatch.setenv("PAIRALIGN_NO_CTA", "1")
    gpu.init()
    gpu.upload(enc)
    got = gpu.align_all_pairs()
    t = gpu.timing()
    assert t["dp_duo_ms"] > 0 and t["dp_cta_ms"] == 0.0
    assert got.tobytes() == want.tobytes()
    monkeypatch.delenv("PAIRALIGN_NO_CTA")
    gpu.init()
    gpu.upload(enc)
    ia, ib = np.array([0, 0, 3]), np.array([1, 2, 4])
    lens = np.array([len(e) for e in enc])
    ops, off, n_ops, res = gpu.align_pairs_ops(ia, ib, lens)
    for k in range(len(ia)):
        r, w = oracle.align_ops(enc[ia[k]], enc[ib[k]], compact=True)
        assert tuple(res[k]) == tuple(r), (ia[k], ib[k])
        assert ops[int(off[k]):int(off[k]) + int(n_ops[k])].tobytes() == w.tobytes(), (ia[k], ib[k])
