"""ctypes binding of include/pairalign_b200.h.

This is plumbing for the tests and bench.py: the product is the shared library
(and the C++ command line built on it).  The library is loaded from the tree
(phylommand_b200/lib/) and loading fails loudly if it has not been built; there is
no Python or CPU implementation to fall back to.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

#: PAIRALIGN_B200_LIB points the binding at another build of the same module (kernel A/B experiments, tools/gpu_run.sh)
LIB_PATH = Path(os.environ.get("PAIRALIGN_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libpairalign_b200.so")

PA_OK, PA_EINVAL, PA_ENODEVICE, PA_ECUDA, PA_ENOMEM, PA_ERANGE = 0, -1, -2, -3, -4, -5

#: numpy view of pa_pair_result (20 bytes, no padding)
RESULT_DTYPE = np.dtype([("score", "<i4"), ("dist", "<u4"), ("len", "<u4"), ("end_i", "<i4"), ("end_j", "<i4")])


class PaParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32),
                ("gap_ext", C.c_int32), ("aligned", C.c_int32)]


class PaTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("total_ms", C.c_double), ("cells", C.c_uint64), ("pairs", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("n_devices", C.c_uint32),
                ("dp_fast_ms", C.c_double), ("dp_general_ms", C.c_double), ("dp_duo_ms", C.c_double), ("dp_cta_ms", C.c_double),
                ("walk_ms", C.c_double)]


#: every symbol include/pairalign_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pa_init": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "pa_shutdown": (None, []),
    "pa_device_count": (C.c_int, []),
    "pa_visible_devices": (C.c_int, []),
    "pa_api_version": (C.c_int, []),
    "pa_last_error": (C.c_char_p, []),
    "pa_char_to_mask": (C.c_int, [C.c_ubyte]),
    "pa_mask_to_char": (C.c_char, [C.c_uint8]),
    "pa_encode_sequence": (C.c_size_t, [C.c_char_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]),
    "pa_upload_sequences": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "pa_num_sequences": (C.c_uint32, []),
    "pa_num_pairs": (C.c_uint64, []),
    "pa_pair_from_index": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pa_align_all_pairs": (C.c_int, [C.POINTER(PaParams), C.c_uint64, C.c_uint64, C.c_void_p]),
    "pa_align_pairs": (C.c_int, [C.POINTER(PaParams), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pa_align_all_pairs_device": (C.c_int, [C.POINTER(PaParams), C.c_uint64, C.c_uint64, C.c_void_p]),
    "pa_align_pair_traceback": (C.c_int, [C.POINTER(PaParams), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]),
    "pa_align_pairs_ops": (C.c_int, [C.POINTER(PaParams), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "pa_partition_pairs": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "pa_partition_by_length": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "pa_count_cells": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "pa_s16_limits": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "pa_get_timing": (C.c_int, [C.POINTER(PaTiming)]),
    "pa_similarity": (C.c_double, [C.c_uint32, C.c_uint32]),
    "pa_pdistance": (C.c_double, [C.c_uint32, C.c_uint32]),
    "pa_jc_distance": (C.c_double, [C.c_uint32, C.c_uint32]),
    "pa_jc_minus_p": (C.c_double, [C.c_uint32, C.c_uint32]),
    "pa_int32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pa_nj_build": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                              C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pa_nj_last_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_lib = None


class PairalignError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pairalign_b200 error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Load the in-tree CUDA module; raises if it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m phylommand_b200.build` "
                               "(the hot path is CUDA only; there is no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != PA_OK:
        raise PairalignError(rc, load().pa_last_error().decode("utf-8", "replace"))


DEFAULT_PARAMS = dict(match=7, mismatch=-5, gap_open=-15, gap_ext=-1, aligned=0)


def make_params(**kw) -> PaParams:
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    return PaParams(**d)


def init(devices=None) -> None:
    lib = load()
    if devices is None:
        _check(lib.pa_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*devices)
        _check(lib.pa_init(arr, len(devices)))


def shutdown() -> None:
    load().pa_shutdown()


def encode(text: bytes | str) -> np.ndarray:
    """translate_to_binary: drops text[0], skips white space / unknown characters."""
    if isinstance(text, str):
        text = text.encode("latin-1")
    out = np.empty(max(len(text), 1), dtype=np.uint8)
    n = load().pa_encode_sequence(text, len(text), out.ctypes.data, None, None, 0)
    return out[:n].copy()


def decode(masks) -> str:
    lib = load()
    return b"".join(lib.pa_mask_to_char(int(m)) for m in masks).decode("ascii")


def pack(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of uint8 mask arrays -> (concatenated masks, uint64 offsets)."""
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        offsets[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    masks = np.concatenate(seqs).astype(np.uint8) if len(seqs) and offsets[-1] > 0 else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(masks), offsets


def upload(seqs) -> None:
    masks, offsets = pack(seqs)
    upload_packed(masks, offsets)


def upload_packed(masks: np.ndarray, offsets: np.ndarray) -> None:
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    _check(load().pa_upload_sequences(masks.ctypes.data if masks.size else None, offsets.ctypes.data, len(offsets) - 1))


def num_pairs() -> int:
    return int(load().pa_num_pairs())


def align_all_pairs(first: int = 0, count: int | None = None, out: np.ndarray | None = None, **params) -> np.ndarray:
    lib = load()
    if count is None:
        count = num_pairs() - first
    if out is None:
        out = np.empty(count, dtype=RESULT_DTYPE)
    assert out.dtype == RESULT_DTYPE and out.flags.c_contiguous and len(out) >= count
    p = make_params(**params)
    _check(lib.pa_align_all_pairs(C.byref(p), first, count, out.ctypes.data))
    return out[:count]


def align_all_pairs_device(d_out_ptr: int, first: int, count: int, **params) -> None:
    p = make_params(**params)
    _check(load().pa_align_all_pairs_device(C.byref(p), first, count, C.c_void_p(d_out_ptr)))


def align_pairs(ia, ib, **params) -> np.ndarray:
    ia = np.ascontiguousarray(ia, dtype=np.uint32)
    ib = np.ascontiguousarray(ib, dtype=np.uint32)
    assert ia.shape == ib.shape
    out = np.empty(len(ia), dtype=RESULT_DTYPE)
    p = make_params(**params)
    _check(load().pa_align_pairs(C.byref(p), ia.ctypes.data, ib.ctypes.data, len(ia), out.ctypes.data))
    return out


def align_pair_traceback(a: int, b: int, cap: int, **params):
    ax = np.empty(cap, dtype=np.uint8)
    ay = np.empty(cap, dtype=np.uint8)
    alen = C.c_uint32(0)
    res = np.empty(1, dtype=RESULT_DTYPE)
    p = make_params(**params)
    _check(load().pa_align_pair_traceback(C.byref(p), a, b, ax.ctypes.data, ay.ctypes.data, cap, C.byref(alen), res.ctypes.data))
    return ax[:alen.value].copy(), ay[:alen.value].copy(), res[0]


def align_pairs_ops(ia, ib, lengths, **params):
    """pairalign -a for a list of pairs: (ops, offsets, n_ops, records); ops[offsets[k]:offsets[k]+n_ops[k]] is pair k."""
    ia = np.ascontiguousarray(ia, dtype=np.uint32)
    ib = np.ascontiguousarray(ib, dtype=np.uint32)
    lengths = np.asarray(lengths, dtype=np.int64)
    cap = int((lengths[ia] + lengths[ib]).sum())
    ops = np.empty(max(cap, 1), dtype=np.uint8)
    offsets = np.empty(len(ia) + 1, dtype=np.uint64)
    n_ops = np.empty(len(ia), dtype=np.uint32)
    res = np.empty(len(ia), dtype=RESULT_DTYPE)
    p = make_params(**params)
    _check(load().pa_align_pairs_ops(C.byref(p), ia.ctypes.data, ib.ctypes.data, len(ia), ops.ctypes.data, cap,
                                     offsets.ctypes.data, n_ops.ctypes.data, res.ctypes.data))
    return ops, offsets, n_ops, res


def s16_limits(**params):
    """(longest sequence with plain 16-bit scores, storage bias) of the s16x2 kernel for these scoring parameters."""
    p = make_params(**params)
    max_len, bias = C.c_uint32(0), C.c_int32(0)
    _check(load().pa_s16_limits(C.byref(p), C.byref(max_len), C.byref(bias)))
    return max_len.value, bias.value


def partition_pairs(first: int, count: int, n_parts: int) -> np.ndarray:
    bounds = np.empty(n_parts + 1, dtype=np.uint64)
    _check(load().pa_partition_pairs(first, count, n_parts, bounds.ctypes.data))
    return bounds


def partition_by_length(lengths, first: int, count: int, n_parts: int):
    """Device-free split of the triangle range into n_parts ranges of nearly equal DP cells."""
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    bounds = np.empty(n_parts + 1, dtype=np.uint64)
    cells = np.empty(n_parts, dtype=np.uint64)
    _check(load().pa_partition_by_length(lengths.ctypes.data, len(lengths), first, count, n_parts,
                                         bounds.ctypes.data, cells.ctypes.data))
    return bounds, cells


def count_cells(first: int, count: int) -> int:
    return int(load().pa_count_cells(first, count))


def pair_from_index(k: int) -> tuple[int, int]:
    a, b = C.c_uint32(), C.c_uint32()
    _check(load().pa_pair_from_index(k, C.byref(a), C.byref(b)))
    return a.value, b.value


def timing() -> dict:
    t = PaTiming()
    _check(load().pa_get_timing(C.byref(t)))
    return {name: getattr(t, name) for name, _ in PaTiming._fields_}


PEAK_CLASSES = {0: "IADD3", 1: "VIMNMX3", 2: "VIADDMNMX", 3: "IMAD", 4: "PRMT", 5: "ISETP+SEL", 6: "IADD3+IMAD mix",
                7: "VIMNMX3.S16x2", 8: "VIADDMNMX.S16x2", 9: "VIADD.16x2", 10: "LOP3", 11: "SHFL", 12: "VIMNMX",
                13: "VIADDMNMX+IMAD mix",
                18: "VIADDMNMX+IMAD imm mix", 19: "VIADDMNMX x3 + IMAD imm", 20: "PRMT+LOP3",
                21: "VIADDMNMX.S16x2+IMAD imm mix"}


def int32_peak(which: int) -> tuple[float, float]:
    g, mhz = C.c_double(), C.c_double()
    _check(load().pa_int32_peak(which, C.byref(g), C.byref(mhz)))
    return g.value, mhz.value


def similarity(dist: int, length: int) -> float:
    return load().pa_similarity(dist, length)


def pdistance(dist: int, length: int) -> float:
    return load().pa_pdistance(dist, length)


def jc_distance(dist: int, length: int) -> float:
    return load().pa_jc_distance(dist, length)


def jc_minus_p(dist: int, length: int) -> float:
    return load().pa_jc_minus_p(dist, length)


NJ_JOIN_DTYPE = np.dtype([("left", "<u4"), ("right", "<u4"), ("left_len", "<f8"), ("right_len", "<f8")])


def nj_build(tri) -> dict:
    """Neighbour joining of a row-major upper-triangle float32 distance matrix (njtree::build_nj_tree).
    Returns joins (NJ_JOIN_DTYPE, n-2), root_left, root_right, root_right_len, kernel_ms, launches, bytes."""
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    n = int(round((1 + (1 + 8 * len(tri)) ** 0.5) / 2))
    assert n * (n - 1) // 2 == len(tri), "not a triangle"
    joins = np.zeros(max(n - 2, 0), dtype=NJ_JOIN_DTYPE)
    rl, rr = C.c_uint32(), C.c_uint32()
    rlen, ms = C.c_double(), C.c_double()
    _check(load().pa_nj_build(tri.ctypes.data, n, joins.ctypes.data if len(joins) else None, C.byref(rl), C.byref(rr),
                              C.byref(rlen), C.byref(ms)))
    launches, nbytes = C.c_uint64(), C.c_uint64()
    load().pa_nj_last_stats(C.byref(launches), C.byref(nbytes))
    return dict(joins=joins, root_left=rl.value, root_right=rr.value, root_right_len=rlen.value, kernel_ms=ms.value,
                launches=launches.value, bytes=nbytes.value)
