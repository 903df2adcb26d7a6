"""Synthetic inputs for the BASELINE.json configurations.

All generators are deterministic in `seed` (numpy PCG64).  Sequences are returned
WITHOUT the sacrificial leading base; `write_fasta` adds one, because the reference
drops the first character of every sequence (src/seqpair.cpp:80).  Names are
s%07d so that std::map order equals generation order.
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
#: ASCII -> 4-bit IUPAC set (A=1 G=2 C=4 T=8), mirrors pa_char_to_mask for the 4 bases
_MASK_OF = np.zeros(256, dtype=np.uint8)
for _ch, _m in zip(b"ACGT", (1, 4, 2, 8)):
    _MASK_OF[_ch] = _m
# IUPAC ambiguity codes (A=1 G=2 C=4 T=8, reference src/seqpair.cpp:22-60)
for _ch, _m in zip(b"RYSWKMBDHVN", (1 | 2, 4 | 8, 2 | 4, 1 | 8, 2 | 8, 1 | 4, 2 | 4 | 8, 1 | 2 | 8, 1 | 4 | 8, 1 | 2 | 4, 15)):
    _MASK_OF[_ch] = _m


def _mutate(seq: np.ndarray, divergence: float, rng: np.random.Generator) -> np.ndarray:
    """80 % substitutions, 10 % deletions of 1-3 bp, 10 % insertions of 1-3 bp."""
    n_events = int(round(divergence * len(seq)))
    if n_events == 0:
        return seq.copy()
    out = seq.tolist()
    for _ in range(n_events):
        kind = rng.random()
        pos = int(rng.integers(0, max(len(out), 1)))
        if kind < 0.8 and out:
            cur = out[pos]
            choices = [b for b in BASES.tolist() if b != cur]
            out[pos] = choices[int(rng.integers(0, 3))]
        elif kind < 0.9 and len(out) > 8:
            k = int(rng.integers(1, 4))
            del out[pos:pos + k]
        else:
            k = int(rng.integers(1, 4))
            out[pos:pos] = BASES[rng.integers(0, 4, size=k)].tolist()
    return np.asarray(out, dtype=np.uint8)


def make_16s_like(n: int, seed: int, root_len: int = 1500, clade_size: int = 40,
                  clade_div: float = 0.08, member_div: float = 0.04):
    """Configs 2/3: one root, clades at `clade_div` from it, members 0..member_div from the clade centre."""
    rng = np.random.default_rng(seed)
    root = BASES[rng.integers(0, 4, size=root_len)]
    n_clades = max(1, (n + clade_size - 1) // clade_size)
    centres = [_mutate(root, clade_div, rng) for _ in range(n_clades)]
    seqs = []
    for s in range(n):
        c = centres[s % n_clades]
        seqs.append(_mutate(c, float(rng.random()) * member_div, rng))
    names = [f"s{s:07d}" for s in range(n)]
    return names, seqs


def make_its_like(n: int, seed: int, n_genera: int = 100):
    """Config 4: variable length 400-900 bp, genus divergence 15-30 %, within-genus 0-6 %.
    Returns names, sequences and taxon strings ('Eukaryota; Fungi; Fam<k>; Gen<g>')."""
    rng = np.random.default_rng(seed)
    n_genera = max(1, min(n_genera, n))
    fam_of = rng.integers(0, max(1, n_genera // 5), size=n_genera)
    roots = []
    fam_root = {}
    for g in range(n_genera):
        f = int(fam_of[g])
        if f not in fam_root:
            L = int(rng.integers(400, 901))
            fam_root[f] = BASES[rng.integers(0, 4, size=L)]
        roots.append(_mutate(fam_root[f], 0.15 + 0.15 * float(rng.random()), rng))
    seqs, names, taxa = [], [], []
    for s in range(n):
        g = int(rng.integers(0, n_genera))
        seqs.append(_mutate(roots[g], 0.06 * float(rng.random()), rng))
        names.append(f"s{s:07d}")
        taxa.append(f"Eukaryota; Fungi; Fam{int(fam_of[g])}; Gen{g}")
    return names, seqs, taxa


def make_long(n: int, seed: int, length: int = 30000, spread: float = 0.05, div_lo: float = 0.03, div_hi: float = 0.15):
    """Config 5: long sequences, length +-5 %, divergence 3-15 % from one root."""
    rng = np.random.default_rng(seed)
    root = BASES[rng.integers(0, 4, size=int(length * (1 + spread)))]
    seqs = []
    for _ in range(n):
        L = int(length * (1 - spread + 2 * spread * float(rng.random())))
        seqs.append(_mutate(root[:L], div_lo + (div_hi - div_lo) * float(rng.random()), rng))
    names = [f"s{s:07d}" for s in range(n)]
    return names, seqs


def make_random(n: int, seed: int, lo: int, hi: int, iupac: float = 0.0, gaps: float = 0.0, related: bool = True):
    """Small mixed test sets: random lengths in [lo,hi], optional IUPAC codes / '-' characters."""
    rng = np.random.default_rng(seed)
    root = BASES[rng.integers(0, 4, size=hi)]
    amb = np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        if related and rng.random() < 0.7:
            s = _mutate(root[:L], 0.3 * float(rng.random()), rng)
        else:
            s = BASES[rng.integers(0, 4, size=L)]
        s = s.copy()
        if len(s) == 0:
            s = BASES[rng.integers(0, 4, size=1)]
        if iupac > 0:
            k = rng.random(len(s)) < iupac
            s[k] = amb[rng.integers(0, len(amb), size=int(k.sum()))]
        if gaps > 0:
            k = rng.random(len(s)) < gaps
            s[k] = ord("-")
        seqs.append(s)
    names = [f"s{s:07d}" for s in range(n)]
    return names, seqs


def to_masks(seq: np.ndarray) -> np.ndarray:
    """Upper-case A/C/G/T/IUPAC byte array -> 4-bit sets (no gaps)."""
    return _MASK_OF[seq]


def to_text(seq: np.ndarray) -> str:
    return seq.tobytes().decode("ascii")


def write_fasta(path, names, seqs, taxa=None, lead: str = "N", width: int = 60) -> None:
    """FASTA with one extra leading base per sequence (the reference drops it)."""
    with open(path, "w") as fh:
        for k, (nm, s) in enumerate(zip(names, seqs)):
            fh.write(f">{nm}" + (f" | {taxa[k]}" if taxa is not None else "") + "\n")
            txt = lead + to_text(s)
            for o in range(0, len(txt), width):
                fh.write(txt[o:o + width] + "\n")
