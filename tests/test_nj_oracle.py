"""Neighbour joining: the C restatement (oracle/nj_oracle.c) against the output of the
UNMODIFIED reference `treeator -n` committed under tests/golden/nj/ (made by oracle/make_nj_golden.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests import oracle_lib

GOLD = Path(__file__).parent / "golden" / "nj"
CASES = json.loads((GOLD / "manifest.json").read_text())


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.load()


@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_oracle_matches_reference_treeator(oracle, case):
    data = (GOLD / f"{case['tag']}.matrix").read_bytes()
    labels = "-L" not in case["flags"]
    parsed = oracle_lib.read_distance_matrix(data, labels)
    assert parsed is not None
    names, tri = parsed
    res = oracle.nj_build(tri)
    got = oracle_lib.newick(names, res, branch_lengths="-0" not in case["flags"])
    assert got == (GOLD / f"{case['tag']}.newick").read_text()


def test_reader_rejects_ragged_matrix():
    assert oracle_lib.read_distance_matrix(b"a 1 2\nb 3 4\nc\n") is None


def test_two_taxa(oracle):
    res = oracle.nj_build(np.array([0.25], dtype=np.float32))
    assert oracle_lib.newick(["a", "b"], res) == "(a:0.000000,b:0.250000);\n"
