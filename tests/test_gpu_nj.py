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
    # bit-exact doubles; a NaN must be a NaN, but its sign and payload are the processor's (x86 makes 0xFFF8.. out
    # of inf-inf, the GPU 0x7FFF..), which only shows as "nan" / "-nan" in a branch length that is meaningless anyway
    for f in ("left_len", "right_len"):
        g, w = got["joins"][f], want["joins"][f]
        assert np.array_equal(np.isnan(g), np.isnan(w))
        ok = ~np.isnan(w)
        assert g[ok].tobytes() == w[ok].tobytes()
    assert (got["root_left"], got["root_right"]) == (want["root_left"], want["root_right"])
    g, w = np.float64(got["root_right_len"]), np.float64(want["root_right_len"])
    assert (np.isnan(g) and np.isnan(w)) or g.tobytes() == w.tobytes()


@pytest.mark.parametrize("kind,n", [("rand", 2), ("rand", 3), ("rand", 4), ("ties", 9), ("rand", 33), ("ties", 257),
                                    ("tree", 300), ("big", 40), ("jc", 200), ("rand", 1000), ("ties", 1500),
                                    ("tree", 2000)])
def test_nj_matches_oracle(gpu, oracle, kind, n):
    tri = _matrix(kind, n, 100 + n)
    got = gpu.nj_build(tri)
    _same(got, oracle.nj_build(tri))
    assert got["launches"] >= 2 * max(n - 2, 0) + 1


def test_nj_epochs_and_wide_matrix(gpu, oracle):
    """3 000 taxa: ~90 compactions, every tile position of the joined taxa and of the new node's column."""
    tri = _matrix("ties", 3000, 3)
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
