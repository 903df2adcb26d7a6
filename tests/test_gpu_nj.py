"""Neighbour joining on the GPU (csrc/pa_nj.cu) through the C-ABI: every join, both
branch lengths (bit-for-bit doubles) and the root against the oracle, and the treeator_b200 -n command line
byte-for-byte against what the unmodified reference treeator printed (tests/golden/nj/)."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import oracle_lib

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "nj"
CASES = json.loads((GOLD / "manifest.json").read_text())


def _matrix(kind: str, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "tree":
        pts = rng.random((n, 6))
        d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)) + 0.05 * rng.random((n, n))
    elif kind == "ties":                      # many exactly equal Q values: the first one in row-major order must win
        d = rng.integers(1, 5, size=(n, n)).astype(float) / 4
    elif kind == "big":                       # nothing below 100000: the reference joins taxa 0 and 1 every round
        d = -(1e7 + rng.random((n, n)))
    elif kind == "jc":                        # what pairalign -j prints: 6 significant digits, some inf / nan / -0
        d = np.round(rng.random((n, n)) * 0.5, 6)
        d[rng.random((n, n)) < 0.01] = np.inf
        d[rng.random((n, n)) < 0.005] = np.nan
        d[rng.random((n, n)) < 0.02] = -0.0
    else:
        d = rng.random((n, n)) * 2
    iu = np.triu_indices(n, 1)
    return d[iu].astype(np.float32)


def _same(got: dict, want: dict):
    assert got["joins"]["left"].tolist() == want["joins"]["left"].tolist()
    assert got["joins"]["right"].tolist() == want["joins"]["right"].tolist()
    # bit-exact doubles (NaN-safe)
    assert got["joins"]["left_len"].tobytes() == want["joins"]["left_len"].tobytes()
    assert got["joins"]["right_len"].tobytes() == want["joins"]["right_len"].tobytes()
    assert (got["root_left"], got["root_right"]) == (want["root_left"], want["root_right"])
    assert np.float64(got["root_right_len"]).tobytes() == np.float64(want["root_right_len"]).tobytes()


@pytest.mark.parametrize("kind,n", [("rand", 2), ("rand", 3), ("rand", 4), ("ties", 9), ("rand", 33), ("ties", 257),
                                    ("tree", 300), ("big", 40), ("jc", 200), ("rand", 1000), ("ties", 1500),
                                    ("tree", 2000)])
def test_nj_matches_oracle(gpu, oracle, kind, n):
    tri = _matrix(kind, n, 100 + n)
    got = gpu.nj_build(tri)
    _same(got, oracle.nj_build(tri))
    assert got["launches"] == 2 * max(n - 2, 0) + 1


@pytest.mark.parametrize("cols", [8, 16, 32])
def test_nj_column_block_variants(gpu, oracle, cols, monkeypatch):
    """The join kernel is instantiated for 8, 16 and 32 columns per CTA (chosen by matrix size): force each."""
    monkeypatch.setenv("PAIRALIGN_NJ_COLS", str(cols))
    for kind, n in (("ties", 130), ("tree", 777)):
        tri = _matrix(kind, n, cols + n)
        _same(gpu.nj_build(tri), oracle.nj_build(tri))


def test_nj_rejects_bad_arguments(gpu):
    from phylommand_b200 import capi
    with pytest.raises(capi.PairalignError):
        gpu.nj_build(np.zeros(0, dtype=np.float32))


@pytest.fixture(scope="module")
def exe():
    from phylommand_b200 import build
    build.build_library()
    return build.build_nj_cli()


@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_treeator_cli_matches_reference(exe, case):
    r = subprocess.run([str(exe), *case["flags"], f"{case['tag']}.matrix"], cwd=GOLD, capture_output=True, timeout=600)
    assert r.returncode == case["rc"], r.stderr.decode(errors="replace")[-2000:]
    assert r.stdout == (GOLD / f"{case['tag']}.newick").read_bytes()


def test_treeator_cli_reads_stdin_and_rejects_ragged(exe):
    data = (GOLD / "synth_ties_17.matrix").read_bytes()
    r = subprocess.run([str(exe), "-n"], input=data, capture_output=True, timeout=600)
    assert r.returncode == 0 and r.stdout == (GOLD / "synth_ties_17.newick").read_bytes()
    r = subprocess.run([str(exe), "-n"], input=b"a 1 2\nb 3 4\nc\n", capture_output=True, timeout=600)
    assert r.returncode == 1 and b"Error in distance matrix" in r.stderr
