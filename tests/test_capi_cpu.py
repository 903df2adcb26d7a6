"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the
header declares, refuses to compute without a device (no CPU fallback), and its front-end
helpers agree with the oracle."""
import math
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "pairalign_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed from the header"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pairalign_b200.h but not exported"
    assert declared == set(capi.SYMBOLS), "ctypes table and header disagree"
    assert lib.pa_api_version() == 1


def test_result_record_layout(capi):
    assert capi.RESULT_DTYPE.itemsize == 20


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for machines without one")
    with pytest.raises(capi.PairalignError) as ei:
        capi.init()
    assert ei.value.code == capi.PA_ENODEVICE
    out = np.zeros(1, dtype=capi.RESULT_DTYPE)
    with pytest.raises(capi.PairalignError) as ei:
        capi.align_all_pairs(0, 1, out)
    assert ei.value.code == capi.PA_ENODEVICE


def test_front_end_matches_oracle(capi, oracle):
    lib = capi.load()
    for c in range(256):
        assert lib.pa_char_to_mask(c) == oracle.lib.pa_oracle_char_mask(c), chr(c)
    for m in range(16):
        assert lib.pa_mask_to_char(m) == oracle.lib.pa_oracle_mask_char(m)
    for text in ("", "A", "NACGT", "xacgu?n.-RYKM \t\r\nswbdhv", "N" + "ACGT" * 50):
        assert capi.encode(text).tolist() == oracle.encode(text).tolist()


def test_stats_match_oracle_bitwise(capi, oracle):
    rng = np.random.default_rng(5)
    pairs = [(0, 0), (0, 7), (3, 4), (4, 4), (1, 3), (749, 1000), (750, 1000), (751, 1000)]
    pairs += [(int(d), int(l)) for l in rng.integers(1, 3000, size=200) for d in [rng.integers(0, l + 1)]]
    for d, l in pairs:
        for mine, theirs in ((capi.similarity, oracle.similarity), (capi.pdistance, oracle.pdist),
                             (capi.jc_distance, oracle.jc), (capi.jc_minus_p, oracle.diff)):
            a, b = mine(d, l), theirs(d, l)
            assert (math.isnan(a) and math.isnan(b)) or a.hex() == b.hex(), (d, l)
