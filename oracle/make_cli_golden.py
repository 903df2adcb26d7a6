#!/usr/bin/env python3
"""Generate tests/golden/cli/: stdout (and .alignment_groups files) of the UNMODIFIED
reference pairalign (oracle/_ref/pairalign, built by oracle/Makefile from /root/reference/src)
for a set of small inputs and flag combinations.  Runs only in the authoring container.

    python oracle/make_cli_golden.py          # a few minutes (the reference is ~10 MCUPS)
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from phylommand_b200 import synth  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "pairalign"
CLI = ROOT / "tests" / "golden" / "cli"
INP = CLI / "inputs"
EX = ROOT / "tests" / "golden" / "example_files"


def write_inputs():
    INP.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(4242)
    # 1. mixed: unsorted names, lower case, IUPAC codes, an unknown character, a repeated accession
    _, seqs = synth.make_random(12, 77, 60, 180, iupac=0.02)
    names = ["zeta", "Alpha", "beta", "gamma_1", "delta", "Beta", "eps", "theta9", "iota", "kappa", "beta", "my seq"]
    with open(INP / "mixed.fst", "w") as fh:
        for k, (nm, s) in enumerate(zip(names, seqs)):
            txt = "N" + synth.to_text(s)
            if k % 3 == 0:
                txt = txt.lower()
            if k == 4:
                txt = txt[:30] + "U?" + txt[30:]
            fh.write(f">{nm}\n")
            for o in range(0, len(txt), 50):
                fh.write(txt[o:o + 50] + "\n")
    # 2. pure A/C/G/T, lengths 300-700 (crosses the 512-column pass width), related sequences
    names, seqs = synth.make_16s_like(10, 5, root_len=560, clade_size=5)
    seqs = [s[: int(rng.integers(300, len(s) + 1))] if k % 2 else s for k, s in enumerate(seqs)]
    synth.write_fasta(INP / "pure.fst", names, seqs)
    # 3. taxon strings + near-identical sequences for clustering / MAD
    names, seqs, taxa = synth.make_its_like(16, 9, n_genera=4)
    seqs = [s[:260] for s in seqs]
    seqs[5] = seqs[4].copy()
    seqs[9] = seqs[8].copy(); seqs[9][10] = ord("A") if seqs[9][10] != ord("A") else ord("C")
    seqs[12] = np.concatenate([seqs[4], np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)])
    synth.write_fasta(INP / "taxa.fst", names, seqs, taxa=taxa)
    # 4. already aligned input with gaps, CRLF line ends
    base = synth.to_text(seqs[0][:120])
    with open(INP / "aligned_crlf.fst", "w", newline="") as fh:
        for k in range(6):
            row = list(base)
            for pos in rng.integers(0, 120, size=10):
                row[int(pos)] = "-" if k % 2 else "ACGT"[int(rng.integers(4))]
            fh.write(f">aln{k}\r\nN" + "".join(row) + "\r\n")
    # 5. a taxonomy file for taxa.fst (overrides the headers)
    with open(INP / "taxonomy.txt", "w") as fh:
        fh.write("Life; Left|" + ",".join(names[:8]) + "\n")
        fh.write("Life; Right|" + " ".join(names[8:]) + "\n")
    # 6. one sequence only; empty file
    (INP / "single.fst").write_text(">only\nNACGTACGTAC\n")
    (INP / "empty.fst").write_text("")


RUNS = [
    # (tag, input, flags)
    ("mixed_a_n", "mixed.fst", ["-a", "-n"]),
    ("mixed_a", "mixed.fst", ["-a"]),
    ("mixed_d_n", "mixed.fst", ["-d", "-n"]),
    ("mixed_d_m_n", "mixed.fst", ["-d", "-m", "-n"]),
    ("mixed_j_m", "mixed.fst", ["-j", "-m"]),
    ("mixed_p_n", "mixed.fst", ["-p", "-n"]),
    ("mixed_s_m_n", "mixed.fst", ["-s", "-m", "-n"]),
    ("mixed_i", "mixed.fst", ["-i"]),
    ("mixed_m", "mixed.fst", ["-m"]),
    ("mixed_long_flags", "mixed.fst", ["--jc_distance", "--names", "--matrix"]),
    ("mixed_A_j_n_m", "mixed.fst", ["-A", "-j", "-n", "-m"]),
    ("mixed_A_a_n", "mixed.fst", ["-A", "-a", "-n"]),
    ("pure_j_n_m", "pure.fst", ["-j", "-n", "-m"]),
    ("pure_a_n", "pure.fst", ["-a", "-n"]),
    ("pure_d", "pure.fst", ["-d"]),
    ("aligned_crlf_A_p_n_m", "aligned_crlf.fst", ["-A", "-p", "-n", "-m"]),
    ("aligned_crlf_j_n_m", "aligned_crlf.fst", ["-j", "-n", "-m"]),
    ("aligned_crlf_a_n", "aligned_crlf.fst", ["-a", "-n"]),
    ("taxa_groups", "taxa.fst", ["--group", "alignment_groups"]),
    ("taxa_both_097", "taxa.fst", ["-g", "both:cut-off=0.97"]),
    ("taxa_both_default", "taxa.fst", ["-g", "both"]),
    ("taxa_both_080", "taxa.fst", ["-g", "both:cut-off=0.80"]),
    ("taxa_cluster_097", "taxa.fst", ["-g", "cluster:cut-off=0.97"]),
    ("taxa_both_taxfile", "taxa.fst", ["-g", "both:cut-off=0.9:taxonomy=taxonomy.txt"]),
    ("taxa_groups_m_n", "taxa.fst", ["-g", "alignment_groups", "-m", "-n"]),
    ("single_j_n_m", "single.fst", ["-j", "-n", "-m"]),
    ("single_d", "single.fst", ["-d", "-n"]),
    ("empty_j_m", "empty.fst", ["-j", "-m"]),
    ("help", None, ["-h"]),
    ("example_taxon_groups", "../../example_files/alignment_file_with_taxon_string.fst", ["--group", "alignment_groups"]),
    ("example_taxon_both_097", "../../example_files/alignment_file_with_taxon_string.fst", ["--group", "both:cut-off=0.97"]),
    ("example_taxon_cluster_097", "../../example_files/alignment_file_with_taxon_string.fst", ["--group", "cluster:cut-off=0.97"]),
]


def main():
    if not REF.exists():
        sys.exit("oracle/_ref/pairalign missing: run `make -C oracle` where /root/reference exists")
    write_inputs()
    manifest = []
    for tag, inp, flags in RUNS:
        cmd = [str(REF), *flags] + ([inp] if inp else [])
        r = subprocess.run(cmd, cwd=INP, capture_output=True)
        (CLI / f"{tag}.out").write_bytes(r.stdout)
        entry = dict(tag=tag, input=inp, flags=flags, rc=r.returncode)
        if inp:
            g = (INP / (inp + ".alignment_groups")).resolve()
            if g.exists():
                (CLI / f"{tag}.alignment_groups").write_bytes(g.read_bytes())
                entry["alignment_groups"] = True
                g.unlink()
        manifest.append(entry)
        print(tag, r.returncode, len(r.stdout), "bytes")
    # pair-fasta round trip: feed the -a -n output back in
    pf = INP / "pairs.pairfst"
    pf.write_bytes((CLI / "pure_a_n.out").read_bytes())
    for tag, flags in (("pairfst_d_n", ["--format", "pairfst", "-d", "-n"]), ("pairfst_A_j", ["--format", "pairfst", "-A", "-j"]),
                       ("pairfst_j_m_n", ["--format", "pairfst", "-j", "-m", "-n"]), ("pairfst_groups", ["--format", "pairfa", "-g", "both:cut-off=0.9"])):
        r = subprocess.run([str(REF), *flags, "pairs.pairfst"], cwd=INP, capture_output=True)
        (CLI / f"{tag}.out").write_bytes(r.stdout)
        entry = dict(tag=tag, input="pairs.pairfst", flags=flags, rc=r.returncode)
        g = INP / "pairs.pairfst.alignment_groups"
        if g.exists():
            (CLI / f"{tag}.alignment_groups").write_bytes(g.read_bytes())
            entry["alignment_groups"] = True
            g.unlink()
        manifest.append(entry)
        print(tag, r.returncode, len(r.stdout), "bytes")
    # the example runs made separately (minutes each): -A, as-is (gaps) and degapped JC matrices
    for tag, inp, flags in (("example_A_jnm", "../../example_files/alignment_file.fst", ["-A", "-j", "-n", "-m"]),
                            ("example_G_jnm", "../../example_files/alignment_file.fst", ["-j", "-n", "-m"]),
                            ("example_D_jnm", "../../example_files/alignment_file_degapped.fst", ["-j", "-n", "-m"])):
        if (CLI / f"{tag}.out").exists():
            manifest.append(dict(tag=tag, input=inp, flags=flags, rc=0))
    (CLI / "manifest.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
