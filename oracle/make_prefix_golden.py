#!/usr/bin/env python3
"""Generate tests/golden/prefix/: the UNMODIFIED reference pairalign (oracle/_ref/pairalign, default build) on
prefixes of the BASELINE.json configurations, as SURVEY.md section 8(d) prescribes for sets the reference cannot
finish: the first 64 sequences of config 3 (10,000 x 1.5 kb, `-j -n -m`) and the first 128 of config 4 (5,000 ITS-like
400-900 bp with taxon strings, `--group both:cut-off=0.97`, stdout and the .alignment_groups file).  The inputs are
committed next to the outputs (the prefix of a 10,000-sequence synthetic set is not the 64-sequence set of the same
seed).  Runs only in the authoring container; about 15 minutes of one core each, run side by side.

    python oracle/make_prefix_golden.py
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from phylommand_b200 import synth  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "pairalign"
OUT = ROOT / "tests" / "golden" / "prefix"

CASES = {
    "c3_64": (["-j", "-n", "-m"], False),
    "c4_128": (["--group", "both:cut-off=0.97"], True),
}


def write_inputs():
    OUT.mkdir(parents=True, exist_ok=True)
    names, seqs = synth.make_16s_like(10000, 1003)
    synth.write_fasta(OUT / "c3_64.fst", names[:64], seqs[:64])
    names, seqs, taxa = synth.make_its_like(5000, 1004)
    synth.write_fasta(OUT / "c4_128.fst", names[:128], seqs[:128], taxa=taxa[:128])


def main():
    if not REF.exists():
        sys.exit("oracle/_ref/pairalign missing: run `make -C oracle` where /root/reference exists")
    write_inputs()
    procs = {}
    for tag, (flags, _) in CASES.items():
        procs[tag] = subprocess.Popen([str(REF), *flags, f"{tag}.fst"], cwd=OUT, stdout=open(OUT / f"{tag}.out", "wb"),
                                      stderr=subprocess.DEVNULL)
    for tag, p in procs.items():
        rc = p.wait()
        print(tag, "rc", rc, (OUT / f"{tag}.out").stat().st_size, "bytes")
        if CASES[tag][1]:
            g = OUT / f"{tag}.fst.alignment_groups"
            g.rename(OUT / f"{tag}.alignment_groups")


if __name__ == "__main__":
    main()
