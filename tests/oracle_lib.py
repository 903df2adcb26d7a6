"""ctypes binding of oracle/liboracle.so and oracle/_ref/libref_seqpair.so.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU baseline
may use this module.  Nothing here reads /root/reference at run time; the oracle is
built by `make -C oracle` (phylommand_b200.build.build_oracle).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libref_seqpair.so"
REF_CLI = ROOT / "oracle" / "_ref" / "pairalign"
REF_CLI_PTHREAD = ROOT / "oracle" / "_ref" / "pairalign_pthread"

REF_TREEATOR = ROOT / "oracle" / "_ref" / "treeator"
NJ_JOIN_DTYPE = np.dtype([("left", "<u4"), ("right", "<u4"), ("left_len", "<f8"), ("right_len", "<f8")])

RESULT_DTYPE = np.dtype([("score", "<i4"), ("dist", "<u4"), ("len", "<u4"), ("end_i", "<i4"), ("end_j", "<i4")])


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.pa_oracle_char_mask.restype = C.c_int
        lib.pa_oracle_char_mask.argtypes = [C.c_ubyte]
        lib.pa_oracle_encode.restype = C.c_size_t
        lib.pa_oracle_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        lib.pa_oracle_mask_char.restype = C.c_char
        lib.pa_oracle_mask_char.argtypes = [C.c_uint8]
        for name in ("pa_oracle_align_full",):
            f = getattr(lib, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        for name in ("pa_oracle_align_ops", "pa_oracle_align_ops_compact"):
            f = getattr(lib, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                          C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        lib.pa_oracle_align_forward.restype = C.c_int
        lib.pa_oracle_align_forward.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_void_p]
        lib.pa_oracle_aligned_stats.restype = None
        lib.pa_oracle_aligned_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        for name in ("pa_oracle_similarity", "pa_oracle_pdist", "pa_oracle_jc", "pa_oracle_diff"):
            f = getattr(lib, name)
            f.restype = C.c_double
            f.argtypes = [C.c_uint32, C.c_uint32]
        lib.pa_oracle_all_pairs.restype = C.c_int
        lib.pa_oracle_all_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]

        lib.nj_oracle_build.restype = C.c_int
        lib.nj_oracle_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_double)]

    def nj_build(self, tri) -> dict:
        """njtree::build_nj_tree restated (oracle/nj_oracle.c): same result layout as capi.nj_build."""
        tri = np.ascontiguousarray(tri, dtype=np.float32)
        n = int(round((1 + (1 + 8 * len(tri)) ** 0.5) / 2))
        assert n * (n - 1) // 2 == len(tri)
        joins = np.zeros(max(n - 2, 0), dtype=NJ_JOIN_DTYPE)
        rl, rr, rlen = C.c_uint32(), C.c_uint32(), C.c_double()
        rc = self.lib.nj_oracle_build(tri.ctypes.data, n, joins.ctypes.data if len(joins) else None, C.byref(rl),
                                      C.byref(rr), C.byref(rlen))
        assert rc == 0
        return dict(joins=joins, root_left=rl.value, root_right=rr.value, root_right_len=rlen.value)

    def encode(self, text) -> np.ndarray:
        if isinstance(text, str):
            text = text.encode("latin-1")
        out = np.empty(max(len(text), 1), dtype=np.uint8)
        n = self.lib.pa_oracle_encode(text, len(text), out.ctypes.data, None)
        return out[:n].copy()

    def decode(self, masks) -> str:
        return b"".join(self.lib.pa_oracle_mask_char(int(m)) for m in masks).decode("ascii")

    def align_full(self, x, y, match=7, mismatch=-5, go=-15, ge=-1):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        ax = np.empty(len(x) + len(y) + 1, dtype=np.uint8)
        ay = np.empty(len(x) + len(y) + 1, dtype=np.uint8)
        alen = C.c_int32(0)
        rc = self.lib.pa_oracle_align_full(x.ctypes.data, len(x), y.ctypes.data, len(y), match, mismatch, go, ge,
                                           res.ctypes.data, ax.ctypes.data, ay.ctypes.data, C.byref(alen))
        assert rc == 0
        return res[0], ax[:alen.value].copy(), ay[:alen.value].copy()

    def align_ops(self, x, y, match=7, mismatch=-5, go=-15, ge=-1, compact=False):
        """(record, op string): 0 = x over y, 1 = x over a gap, 2 = a gap over y.  compact: 2 bits per cell instead of
        the three full matrices (long pairs)."""
        x = np.ascontiguousarray(x, dtype=np.uint8)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        ops = np.empty(len(x) + len(y) + 1, dtype=np.uint8)
        alen = C.c_int32(0)
        f = self.lib.pa_oracle_align_ops_compact if compact else self.lib.pa_oracle_align_ops
        rc = f(x.ctypes.data, len(x), y.ctypes.data, len(y), match, mismatch, go, ge, res.ctypes.data, ops.ctypes.data, C.byref(alen))
        assert rc == 0
        return res[0], ops[:alen.value].copy()

    def align_forward(self, x, y, match=7, mismatch=-5, go=-15, ge=-1):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        rc = self.lib.pa_oracle_align_forward(x.ctypes.data, len(x), y.ctypes.data, len(y), match, mismatch, go, ge,
                                              res.ctypes.data)
        assert rc == 0
        return res[0]

    def aligned_stats(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        self.lib.pa_oracle_aligned_stats(x.ctypes.data, len(x), y.ctypes.data, len(y), res.ctypes.data)
        return res[0]

    def all_pairs(self, masks, offsets, first=0, last=None, threads=1, match=7, mismatch=-5, go=-15, ge=-1):
        masks = np.ascontiguousarray(masks, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if last is None:
            last = n * (n - 1) // 2
        out = np.zeros(last - first, dtype=RESULT_DTYPE)
        rc = self.lib.pa_oracle_all_pairs(masks.ctypes.data, offsets.ctypes.data, n, match, mismatch, go, ge,
                                          first, last, threads, out.ctypes.data)
        assert rc == 0
        return out

    def similarity(self, d, l): return self.lib.pa_oracle_similarity(d, l)
    def pdist(self, d, l): return self.lib.pa_oracle_pdist(d, l)
    def jc(self, d, l): return self.lib.pa_oracle_jc(d, l)
    def diff(self, d, l): return self.lib.pa_oracle_diff(d, l)


class RefSeqpair:
    """The UNMODIFIED reference seqpair class (oracle/_ref/libref_seqpair.so)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_seqpair_run.restype = C.c_int
        lib.ref_seqpair_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p, C.c_char_p, C.c_int]

    def run(self, x: str, y: str, aligned: bool = False):
        """x, y: raw text as pairalign passes it (first character is dropped by the reference)."""
        cap = len(x) + len(y) + 8
        ax = C.create_string_buffer(cap)
        ay = C.create_string_buffer(cap)
        score, ham = C.c_int(), C.c_int()
        sim, jc = C.c_double(), C.c_double()
        rc = self.lib.ref_seqpair_run(x.encode("latin-1"), y.encode("latin-1"), int(aligned), C.byref(score), C.byref(ham),
                                      C.byref(sim), C.byref(jc), ax, ay, cap)
        assert rc == 0
        return dict(score=score.value, hamming=ham.value, sim=sim.value, jc=jc.value,
                    x=ax.value.decode("ascii"), y=ay.value.decode("ascii"))


def ensure_built() -> None:
    srcs = [ROOT / "oracle" / n for n in ("pa_oracle.c", "pa_oracle.h", "nj_oracle.c")]
    if not ORACLE_SO.exists() or any(s.stat().st_mtime > ORACLE_SO.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), str(ORACLE_SO)], check=True)


def load() -> Oracle:
    ensure_built()
    return Oracle(C.CDLL(str(ORACLE_SO)))


def load_ref():
    """None when the compiled reference is not available (it is built in the authoring container)."""
    if not REF_SO.exists():
        return None
    return RefSeqpair(C.CDLL(str(REF_SO)))


# ---- neighbour joining: text in, text out (test-side restatement of the reference's reader and printer) ----

def read_distance_matrix(data: bytes, labels: bool = True):
    """njtree::read_distance_matrix (src/nj_tree.cpp:252-352) -> (names, float32 upper triangle) or None when
    matrix_good() (src/nj_tree.cpp:22-30) would fail."""
    rows, names = [], []
    new_row, value, n_taxa = True, b"", 0
    for k, ch in enumerate(data):
        ch = bytes([ch])
        at_end = k + 1 == len(data)                   # infile.peek() == EOF: the last character ends a value, unread
        if ch in (b" ", b"\n", b"\r", b"\t") or at_end:
            if value:
                if new_row:
                    rows.append([])
                    if labels:
                        names.append(value.decode("latin-1"))
                    else:
                        names.append(str(n_taxa))
                        rows[-1].append(np.float32(_atof(value)))
                    n_taxa += 1
                    if ch not in (b"\n", b"\r"):
                        new_row = False
                else:
                    rows[-1].append(np.float32(_atof(value)))
                value = b""
        else:
            value += ch
        if ch in (b"\n", b"\r"):
            new_row = True
    if rows and rows[-1]:
        rows.append([])
        names.append(str(n_taxa))
    n = len(rows)
    for k, row in enumerate(rows):
        if len(row) != n - 1 - k:
            return None
    tri = np.array([v for row in rows for v in row], dtype=np.float32)
    return names, tri


def _atof(tok: bytes) -> float:
    import re
    m = re.match(rb"[ \t]*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|inf(inity)?|nan)", tok, re.I)
    return float(m.group(0)) if m else 0.0


def newick(names, res: dict, branch_lengths: bool = True) -> str:
    """tree::print_newick (src/tree.cpp:239-279): labels, ':' << fixed << branchlength on every node but the root."""
    n = len(names)
    joins = res["joins"]
    blen = {}
    for j in joins:
        blen[int(j["left"])] = float(j["left_len"])
        blen[int(j["right"])] = float(j["right_len"])
    blen[res["root_left"]] = 0.0
    blen[res["root_right"]] = float(res["root_right_len"])
    out = ["("]
    # explicit stack: 10 000 taxa would overflow Python's recursion limit
    st = [(res["root_right"], 0), (None, 3), (res["root_left"], 0)]
    while st:
        nid, phase = st.pop()
        if phase == 3:
            out.append(",")
        elif nid < n:
            out.append(names[nid])
            if branch_lengths:
                out.append(":%f" % blen[nid])
        elif phase == 0:
            j = joins[nid - n]
            out.append("(")
            st += [(nid, 2), (int(j["right"]), 0), (None, 3), (int(j["left"]), 0)]
        else:
            out.append(")")
            if branch_lengths:
                out.append(":%f" % blen[nid])
    out.append(");\n")
    return "".join(out)
