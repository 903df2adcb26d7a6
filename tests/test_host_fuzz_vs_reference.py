"""Differential fuzz of the command line's host side against the UNMODIFIED reference binary
(oracle/_ref/pairalign, compiled by oracle/Makefile where /root/reference exists): random small FASTA
files -- unsorted and repeated accessions, names with blanks, taxonomy strings of different depths, lower
case, IUPAC codes, line wrapping, CRLF -- through every output mode, matrix framing, clustering / MAD
runs and pair-FASTA round trips.  build/pairalign_hosttest is the product's host code with the device half
replaced by the oracle-backed double (tests/host_double/), so this also fuzzes the oracle's DP against the
reference's.  Skipped where the reference binary is absent (it travels to the GPU box, the CPU suite there
is not run)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.test_host_replay_cpu import exe  # noqa: F401  (fixture: builds build/pairalign_hosttest)

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "pairalign"

pytestmark = pytest.mark.skipif(not REF.exists(), reason="the compiled reference (oracle/_ref/pairalign) is not here")

BASES = "ACGT"
IUPAC = "RYSWKMBDHVN"
FAMILIES = ["Eukaryota; Fungi; Fam1; GenA", "Eukaryota; Fungi; Fam1; GenB", "Eukaryota; Fungi; Fam2; GenC", "Eukaryota; Fungi; Fam2",
            "Eukaryota; Plantae; FamP; GenP", "Eukaryota"]


def mutate(rng, root, rate):
    out = []
    for c in root:
        u = rng.random()
        if u < rate * 0.7:
            out.append(BASES[int(rng.integers(4))])
        elif u < rate * 0.85:
            continue
        elif u < rate:
            out.append(c)
            out.append(BASES[int(rng.integers(4))])
        else:
            out.append(c)
    return "".join(out)


def make_case(seed, group):
    """One random FASTA text.  group: inputs for --group runs (related sequences with taxonomy: keeps JC finite,
    the reference indexes its histogram with the value and a NaN is undefined behaviour there)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 11))
    L = int(rng.integers(30, 90)) if group else int(rng.integers(4, 70))
    root = "".join(BASES[int(k)] for k in rng.integers(0, 4, size=L))
    names, lines = [], []
    for k in range(n):
        rate = float(rng.random()) * (0.15 if group else 0.5)
        seq = mutate(rng, root, rate)
        if not group and rng.random() < 0.3:
            seq = seq[: max(3, int(rng.integers(3, len(seq) + 1)))]
        seq = list(seq) if len(seq) >= 3 else list(root[:3])
        if not group:
            for pos in range(len(seq)):
                if rng.random() < 0.03:
                    seq[pos] = IUPAC[int(rng.integers(len(IUPAC)))]
        if not group and seed % 3 == 0:                      # gap characters in unaligned input: INT_MIN scores, 32-bit wrap-around
            for pos in range(len(seq)):
                if rng.random() < 0.06:
                    seq[pos] = "-"
        if not group and seed % 5 == 0 and len(seq) > 6:     # characters outside the alphabet: skipped with a warning
            seq.insert(int(rng.integers(1, len(seq))), "U")
            seq.insert(int(rng.integers(1, len(seq))), "?")
        seq = "".join(seq)
        if rng.random() < 0.3:
            seq = seq.lower()
        name = f"s{int(rng.integers(0, 40)):02d}" if rng.random() < 0.8 else f"Seq {k} x"      # repeats and blanks happen
        head = ">" + name
        if group or rng.random() < 0.3:
            if rng.random() < 0.9:
                head += " | " + FAMILIES[int(rng.integers(len(FAMILIES)))]
        names.append(name)
        lines.append(head)
        text = BASES[int(rng.integers(4))] + seq           # the reference drops the first character
        width = int(rng.integers(10, 80))
        lines += [text[o:o + width] for o in range(0, len(text), width)]
    eol = "\r\n" if (not group and seed % 7 == 0) else "\n"
    return eol.join(lines) + eol


MODES = [["-j", "-n", "-m"], ["-d", "-n"], ["-d", "-m"], ["-p", "-m", "-n"], ["-s"], ["-i", "-n", "-m"], ["-a", "-n"], ["-a"],
         ["-j"], ["-A", "-p", "-n", "-m"], ["-A", "-d", "-n"], ["-m", "-a", "-n"], ["-m", "-a"], ["-a", "-v"]]
GROUP_MODES = [["-g", "alignment_groups"], ["-g", "both:cut-off=0.9"], ["-g", "both:cut-off=0.97"], ["-g", "cluster:cut-off=0.9"],
               ["-g", "both"], ["-g", "alignment_groups", "-m", "-n"], ["-g", "both:cut-off=0.8", "-n"]]


def run_both(exe, tmp_path, flags, name):  # noqa: F811
    env = dict(os.environ)
    outs = []
    for binary, sub in ((REF, "ref"), (exe, "ours")):
        d = tmp_path / sub
        d.mkdir(exist_ok=True)
        (d / name).write_bytes((tmp_path / name).read_bytes())
        r = subprocess.run([str(binary), *flags, name], cwd=d, capture_output=True, timeout=300, env=env)
        groups = d / (name + ".alignment_groups")
        outs.append((r.returncode, r.stdout, groups.read_bytes() if groups.exists() else None,
                     r.stderr.count(b"Can not interpret")))
        if groups.exists():
            groups.unlink()
    return outs


@pytest.mark.parametrize("seed", range(40))
def test_output_modes_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    (tmp_path / "in.fst").write_bytes(make_case(1000 + seed, group=False).encode())
    for flags in MODES:
        ref, ours = run_both(exe, tmp_path, flags, "in.fst")
        assert ours[0] == ref[0], (flags, seed)
        assert ours[1] == ref[1], (flags, seed)
        assert ours[3] == ref[3], (flags, seed)          # as many "Can not interpret" warnings as the reference prints


def test_single_sequence_and_two_sequences(exe, tmp_path):  # noqa: F811
    """N = 1: the reference still visits one pair with an empty second sequence; N = 2: one pair, one matrix row."""
    for text in (">only\nNACGTACGT\n", ">b\nNACGTTTGA\n>a\nNACGATTGA\n", ">x | Eukaryota; Fungi\nNACGTACGTAA\n"):
        (tmp_path / "in.fst").write_bytes(text.encode())
        for flags in MODES + [["-g", "alignment_groups"], ["-g", "both:cut-off=0.9"]]:
            if text.count(">") == 1 and flags[0] == "-A":
                continue                                   # -A with one sequence compares against an empty string: position-wise UB
            ref, ours = run_both(exe, tmp_path, flags, "in.fst")
            assert ours[:3] == ref[:3], (flags, text)


@pytest.mark.parametrize("seed", range(24))
def test_group_modes_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    (tmp_path / "in.fst").write_bytes(make_case(2000 + seed, group=True).encode())
    for flags in GROUP_MODES:
        ref, ours = run_both(exe, tmp_path, flags, "in.fst")
        assert ours[0] == ref[0], (flags, seed)
        assert ours[1] == ref[1], (flags, seed)
        assert ours[2] == ref[2], (flags, seed)


@pytest.mark.parametrize("seed", range(8))
def test_pairfasta_round_trip_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    """-a -n output read back with --format pairfst (the miniptera.pl pipeline): the first character is dropped again."""
    (tmp_path / "in.fst").write_bytes(make_case(3000 + seed, group=False).encode())
    ref, ours = run_both(exe, tmp_path, ["-a", "-n"], "in.fst")
    assert ours[1] == ref[1]
    (tmp_path / "pairs.pairfst").write_bytes(ref[1])
    for flags in (["--format", "pairfst", "-A", "-j", "-n"], ["--format", "pairfst", "-d", "-n"], ["--format", "pairfst", "-p", "-m", "-n"],
                  ["--format", "pairfst", "-g", "alignment_groups"]):
        r, o = run_both(exe, tmp_path, flags, "pairs.pairfst")
        assert o[0] == r[0], (flags, seed)
        assert o[1] == r[1], (flags, seed)
        assert o[2] == r[2], (flags, seed)


def make_odd_case(seed):
    """Headers and layout the FASTA index has to treat the reference's way: no accession (numbered), blanks inside
    names (removed), several '|', taxonomy with leading blanks or the literal 'empty', blank lines, tabs and blanks
    inside sequence lines, very short sequences."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 9))
    L = int(rng.integers(12, 60))
    root = "".join(BASES[int(k)] for k in rng.integers(0, 4, size=L))
    lines = []
    for k in range(n):
        seq = mutate(rng, root, float(rng.random()) * 0.2)
        if rng.random() < 0.2:
            seq = seq[: int(rng.integers(1, 4))]                 # one to three bases after the dropped one
        if len(seq) < 1:
            seq = "A"
        style = int(rng.integers(0, 8))
        name = f"n{int(rng.integers(0, 30))}"
        tax = FAMILIES[int(rng.integers(len(FAMILIES)))]
        head = {0: f">{name}", 1: ">", 2: f"> {name} extra words", 3: f">{name}|{tax}", 4: f">{name} |   {tax}",
                5: f">{name} | {tax} | something else", 6: "> | " + tax, 7: f">{name} | empty"}[style]
        lines.append(head)
        text = BASES[int(rng.integers(4))] + seq
        width = int(rng.integers(5, 40))
        for o in range(0, len(text), width):
            piece = text[o:o + width]
            if rng.random() < 0.2 and len(piece) > 2:
                cut = int(rng.integers(1, len(piece)))
                piece = piece[:cut] + (" " if rng.random() < 0.5 else "\t") + piece[cut:]
            lines.append(piece)
        if rng.random() < 0.3:
            lines.append("")
    # always a final newline: without one the reference's stream fails after the first read of the last record
    # and every later sequence comes back empty (undefined behaviour from there on; documented deviation)
    return "\n".join(lines) + "\n"


ODD_MODES = [["-j", "-n", "-m"], ["-d", "-n"], ["-a", "-n"], ["-s", "-m"], ["-A", "-i", "-n"], ["-g", "alignment_groups"],
             ["-g", "both:cut_off=0.93"], ["-g", "both:cutoff=0.5:only_lead"], ["-g", "both:cut-off=all,0.9"],
             ["-g", "both:cut-off=in.fst,0.85"], ["-g", "both:cut-off=other,0.85"], ["-g", "alignment_groups:min_length=20"],
             ["-g", "both:cut-off=0.9:taxonomy=tax.txt"], ["-g", "alignment_groups:taxonomy=tax.txt", "-n"],
             ["-g", "both:cut-off=0.9:taxonomy=missing.txt"]]


@pytest.mark.parametrize("seed", range(30))
def test_odd_headers_and_group_arguments_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    text = make_odd_case(4000 + seed)
    (tmp_path / "in.fst").write_bytes(text.encode())
    # a taxonomy file naming some of the accessions (src/pairalign.cpp:409-439)
    rng = np.random.default_rng(seed)
    rows = []
    for fam in FAMILIES[:4]:
        accs = [f"n{int(k)}" for k in rng.integers(0, 30, size=int(rng.integers(1, 6)))]
        rows.append(fam + "|" + (", " if rng.random() < 0.5 else " ").join(accs))
    for sub in ("ref", "ours"):
        (tmp_path / sub).mkdir(exist_ok=True)
        (tmp_path / sub / "tax.txt").write_text("\n".join(rows) + "\n")
    for flags in ODD_MODES:
        ref, ours = run_both(exe, tmp_path, flags, "in.fst")
        assert ours[0] == ref[0], (flags, seed)
        assert ours[1] == ref[1], (flags, seed)
        assert ours[2] == ref[2], (flags, seed)


# ---- treeator -n: reader, neighbour joining (oracle) and printer against the reference's treeator ------------------------
REF_TREEATOR = ROOT / "oracle" / "_ref" / "treeator"


def make_matrix(seed):
    """A triangular matrix as pairalign -m prints it, with the liberties a hand-made file takes: tabs and runs of
    blanks, CRLF, values in scientific notation, -0 / inf / nan, integer ties, labels alone on a line."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 14))
    style = int(rng.integers(0, 4))
    labels = bool(seed % 4)
    pts = rng.random((n, 3))
    rows = []
    for a in range(n - 1):
        vals = []
        for b in range(a + 1, n):
            d = float(np.linalg.norm(pts[a] - pts[b]))
            if style == 1:
                d = float(int(d * 4))                       # many exact ties
            txt = f"{d:.6g}"
            u = rng.random()
            if style == 2 and u < 0.1:
                txt = f"{d:.3e}"
            elif style == 3 and u < 0.06:
                txt = ["-0", "inf", "nan", "0"][int(rng.integers(4))]
            vals.append(txt)
        sep = "\t" if rng.random() < 0.2 else " " * int(rng.integers(1, 3))
        row = (f"t{a:02d}" + sep if labels else "") + " " * a + sep.join(vals) + " "
        if labels and rng.random() < 0.1:
            row = f"t{a:02d}\n" + " " * a + sep.join(vals) + " "        # the label alone on its line
        rows.append(row)
    eol = "\r\n" if seed % 9 == 0 else "\n"
    text = eol.join(rows) + eol
    if labels or rng.random() < 0.5:
        text += (f"t{n - 1:02d}" if labels else str(n - 1)) + eol      # pairalign prints the last name with or without -n
    return text, labels


@pytest.mark.skipif(not REF_TREEATOR.exists(), reason="the compiled reference treeator is not here")
@pytest.mark.parametrize("seed", range(60))
def test_treeator_host_side_against_the_reference(tmp_path, seed):
    from tests.test_host_replay_cpu import _compile
    from phylommand_b200 import build
    from tests import oracle_lib
    build.build_library()
    oracle_lib.load()
    out = ROOT / "build" / "treeator_hosttest"
    srcs = [ROOT / "phylommand_b200" / "host_nj" / "treeator_nj_main.cpp", ROOT / "tests" / "host_double" / "nj_build_oracle.cpp"]
    _compile(out, srcs, srcs, build.LIB_DIR, [])
    text, labels = make_matrix(7000 + seed)
    (tmp_path / "m.txt").write_bytes(text.encode())
    for flags in (["-n"], ["-n", "-0"]):
        flags = flags + ([] if labels else ["-L"])
        ref = subprocess.run([str(REF_TREEATOR), *flags, "m.txt"], cwd=tmp_path, capture_output=True, timeout=120)
        ours = subprocess.run([str(out), *flags, "m.txt"], cwd=tmp_path, capture_output=True, timeout=120)
        if b"nan" in ref.stdout:                     # x86 prints inf - inf as -nan: the documented sign-of-NaN deviation
            assert ours.stdout.replace(b"-nan", b"nan") == ref.stdout.replace(b"-nan", b"nan"), (flags, seed)
        else:
            assert ours.stdout == ref.stdout, (flags, seed)
        assert ours.returncode == ref.returncode, (flags, seed)


LONG_AND_ERRORS = [["--jc_distance", "--names", "--matrix"], ["--distances", "--names"], ["--proportion_difference", "--matrix"],
                   ["--similarity"], ["--difference", "--names", "--matrix"], ["--alignments", "--names"], ["--aligned", "--jc_distance"],
                   ["--group", "both:cut-off=0.9"], ["--group", "alignment_groups", "--verbose"], ["-j", "-n", "-m", "-v"],
                   ["--format", "fasta", "-j"], ["--format", "pairfa", "-j"], ["-h"], ["--help"], ["-j", "-x"], ["--format", "xml"],
                   ["-g", "nonsense"], ["-g", "both:foo=1"], ["-g"], ["-T", "2", "-j", "-m"], ["-T"], ["--threads", "0"]]


@pytest.mark.parametrize("seed", range(6))
def test_long_options_stdin_and_argument_errors_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    """Long option names, input on stdin (the alignment groups then go to sequence.alignment_groups), help, and
    the argument errors with their exit codes.  -T/--threads is 'not recognized' by the reference's default build (no
    PTHREAD) and therefore here."""
    text = make_case(5000 + seed, group=(seed % 2 == 0))
    for flags in LONG_AND_ERRORS:
        for use_stdin in (False, True):
            outs = []
            for binary, sub in ((REF, "ref"), (exe, "ours")):
                d = tmp_path / f"{sub}_{int(use_stdin)}"
                d.mkdir(exist_ok=True)
                for old in d.iterdir():
                    old.unlink()
                (d / "in.fst").write_text(text)
                if use_stdin:
                    r = subprocess.run([str(binary), *flags], cwd=d, input=text.encode(), capture_output=True, timeout=120)
                else:
                    r = subprocess.run([str(binary), *flags, "in.fst"], cwd=d, capture_output=True, timeout=120)
                made = sorted((p.name, p.read_bytes()) for p in d.iterdir() if p.name != "in.fst")
                outs.append((r.returncode, r.stdout, made))
            assert outs[0] == outs[1], (flags, use_stdin, seed)


@pytest.mark.parametrize("seed", range(6))
def test_verbose_stderr_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    """-v: the progress dots, the messages and the approximate MAD of the whole group on stderr (840 such runs were
    compared once: no difference).  The program path and the two time stamps are masked."""
    import re
    group = seed % 2 == 0
    (tmp_path / "in.fst").write_text(make_case(6000 + seed, group))

    def norm(raw, path):
        text = raw.decode(errors="replace").replace(str(path), "BIN")
        return re.sub(r"(Sat|Sun|Mon|Tue|Wed|Thu|Fri) \w{3} +\d+ [\d:]+ \d{4}", "DATE", text)

    for flags in (GROUP_MODES if group else MODES[:5]) + [["-g", "alignment_groups"]]:
        outs = []
        for binary, sub in ((REF, "ref"), (exe, "ours")):
            d = tmp_path / sub
            d.mkdir(exist_ok=True)
            (d / "in.fst").write_bytes((tmp_path / "in.fst").read_bytes())
            r = subprocess.run([str(binary), *flags, "-v", "in.fst"], cwd=d, capture_output=True, timeout=120)
            outs.append((r.returncode, r.stdout, norm(r.stderr, binary)))
        assert outs[0] == outs[1], (flags, seed)


@pytest.mark.parametrize("seed", range(4))
def test_verbose_stderr_in_pairfst_mode_against_the_reference(exe, tmp_path, seed):  # noqa: F811
    """--format pairfst with -v: the same messages as the reference on stderr ('All sequences will be treated as from
    same taxon.', 'No alignment_groups file/table present...', 'Using the cut off', 'Checking <table>', 'Finished
    aligning...'), with and without a taxonomy file; and the un-silenced 'Could not initiate sequence retrieval' line
    for a pairfst file that cannot be read."""
    import re

    def norm(raw, path):
        text = raw.decode(errors="replace").replace(str(path), "BIN")
        return re.sub(r"(Sat|Sun|Mon|Tue|Wed|Thu|Fri) \w{3} +\d+ [\d:]+ \d{4}", "DATE", text)

    (tmp_path / "in.fst").write_bytes(make_case(8000 + seed, group=False).encode())
    ref, ours = run_both(exe, tmp_path, ["-a", "-n"], "in.fst")
    assert ours[1] == ref[1]
    names = sorted({ln[1:].split(b"|")[0].strip().decode() for ln in ref[1].split(b"\n") if ln.startswith(b">")})
    tax = "Life; A|" + ",".join(names[: len(names) // 2]) + "\nLife; B|" + " ".join(names[len(names) // 2:]) + "\n"
    cases = [(["--format", "pairfst", "-j", "-n"], "pairs.pairfst"), (["--format", "pairfst", "-g", "alignment_groups"], "pairs.pairfst"),
             (["--format", "pairfst", "-g", "both:cut-off=0.9"], "pairs.pairfst"),
             (["--format", "pairfst", "-g", "both:cut-off=0.9:taxonomy=tax.txt"], "pairs.pairfst"),
             (["--format", "pairfst", "-g", "cluster:cut-off=0.95"], "pairs.pairfst"),
             (["--format", "pairfst", "-d", "-m"], "missing.pairfst"), (["--format", "pairfst", "-g", "both"], "missing.pairfst")]
    for flags, name in cases:
        for verbose in (["-v"], []):
            outs = []
            for binary, sub in ((REF, "ref"), (exe, "ours")):
                d = tmp_path / f"{sub}_v"
                d.mkdir(exist_ok=True)
                for old in d.iterdir():
                    old.unlink()
                (d / "pairs.pairfst").write_bytes(ref[1])
                (d / "tax.txt").write_text(tax)
                r = subprocess.run([str(binary), *flags, *verbose, name], cwd=d, capture_output=True, timeout=120)
                made = sorted((p.name, p.read_bytes()) for p in d.iterdir() if p.name not in ("pairs.pairfst", "tax.txt"))
                outs.append((r.returncode, r.stdout, norm(r.stderr, binary), made))
            assert outs[0] == outs[1], (flags, name, verbose, seed)


@pytest.mark.parametrize("seed", range(4))
def test_input_without_final_newline_reads_like_the_same_file_with_one(exe, tmp_path, seed):  # noqa: F811
    """The one deliberate deviation of the FASTA front end (host/fasta_index.h): without a final newline
    the reference duplicates the last character (stdin) or leaves its stream failed and reads empty sequences from
    then on (file).  Here such input gives what the reference gives for the same text WITH the newline."""
    text = make_case(9000 + seed, group=False)
    assert text.endswith("\n")
    for flags in (["-j", "-n", "-m"], ["-a", "-n"], ["-d"]):
        (tmp_path / "in.fst").write_bytes(text.encode())
        ref, _ = run_both(exe, tmp_path, flags, "in.fst")
        d = tmp_path / "cut"
        d.mkdir(exist_ok=True)
        (d / "in.fst").write_bytes(text.rstrip("\r\n").encode())
        from_file = subprocess.run([str(exe), *flags, "in.fst"], cwd=d, capture_output=True, timeout=120)
        from_stdin = subprocess.run([str(exe), *flags], cwd=d, input=text.rstrip("\r\n").encode(), capture_output=True, timeout=120)
        assert from_file.stdout == ref[1] and from_stdin.stdout == ref[1], (flags, seed)
