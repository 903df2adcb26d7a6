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


def test_s16_storage_bias_keeps_every_state_negative_and_in_range(capi):
    """The s16x2 kernel stores states as true value + bias so that H + GO can be a plain 32-bit add.
    True values lie in [-(|ge| L + |mismatch| + |go|), match L] (checked against the literal DP in the test below);
    with the bias both ends must stay negative, and one more gap extension must still fit int16."""
    for sc in (dict(), dict(match=5, mismatch=-4, gap_open=-10, gap_ext=-2), dict(match=20, mismatch=-25, gap_open=-100, gap_ext=-7),
               dict(match=1, mismatch=-1, gap_open=0, gap_ext=0), dict(match=127, mismatch=-1, gap_open=-127, gap_ext=-50)):
        p = dict(capi.DEFAULT_PARAMS); p.update(sc)
        L, bias = capi.s16_limits(**sc)
        assert L >= 16 and bias < 0, sc
        top = max(p["match"], 1) * L + bias
        bottom = bias - (max(-p["gap_ext"], 1) * (L + 1) + abs(p["mismatch"]) + abs(p["gap_open"]))
        assert top <= -16 and bottom >= -32768, (sc, L, bias, top, bottom)
    assert capi.s16_limits() == (4059, -16 - 7 * 4059)
    assert capi.s16_limits(match=1000) == (0, 0)            # outside the byte tables: general kernel
    assert capi.s16_limits(match=127, mismatch=-128, gap_open=-127) == (0, 0)      # mismatch + gap_open does not fit a byte


def test_true_value_range_of_the_recurrence(oracle):
    """The bound the bias rests on, on the literal recurrence: no state (nor H + GO, nor a gap state + GE) ever leaves
    [-(L + 20), 7 L] for pairalign's scoring, including the inputs built to go low (nothing matches) or high (identical)."""
    rng = np.random.default_rng(5)
    M, X, GO, GE = 7, -5, -15, -1
    cases = [(np.zeros(70, int), np.ones(70, int)), (np.zeros(70, int), np.zeros(70, int)), (np.arange(70) % 2, (np.arange(70) + 1) % 2),
             (np.zeros(3, int), np.ones(70, int)), (np.ones(70, int), np.zeros(3, int))]
    cases += [(rng.integers(0, 4, int(rng.integers(1, 60))), rng.integers(0, 2, int(rng.integers(1, 60)))) for _ in range(25)]
    for x, y in cases:
        n, m = len(x), len(y)
        H = np.zeros((n, m), int); Gy = np.zeros((n, m), int); Gx = np.zeros((n, m), int)
        lo, hi = 0, 0
        for i in range(n):
            for j in range(m):
                s = M if x[i] == y[j] else X
                if i == 0 or j == 0:
                    H[i, j] = s
                else:
                    o = H[i - 1, j - 1] + GO
                    u, l = Gy[i - 1, j] + GE, Gx[i, j - 1] + GE
                    H[i, j] = max(H[i - 1, j - 1], Gy[i - 1, j], Gx[i, j - 1]) + s
                    Gy[i, j], Gx[i, j] = max(o, u), max(o, l)
                    lo = min(lo, o, u, l)
                lo, hi = min(lo, H[i, j]), max(hi, H[i, j], Gy[i, j], Gx[i, j])
        L = max(n, m)
        assert lo >= -(L + 20) and hi <= 7 * L, (lo, hi, L)
