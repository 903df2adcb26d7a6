"""The command line's host side on a machine WITHOUT a GPU: build/pairalign_hosttest is the product's own
host code (FASTA index and order, pipeline, matrix framing, number formatting, single-link clusters, MAD
groups, pair-FASTA input, rendering of alignments) linked with a test double of its device half
(tests/host_double/seqpair_batch_oracle.cpp: records and op strings from the oracle).  Its output must be
byte for byte what the reference's pairalign printed (tests/golden/cli/*.out, made by
oracle/make_cli_golden.py from the unmodified reference).  The product binary build/pairalign_b200 never
contains the double; tests/test_gpu_cli.py runs the same command lines through the CUDA module."""
import json
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CLI_DIR = ROOT / "tests" / "golden" / "cli"
MANIFEST = json.loads((CLI_DIR / "manifest.json").read_text())
HOST = ROOT / "phylommand_b200" / "host"


def _compile(out, deps, srcs, lib_dir, extra):
    """g++ into a private name, then an atomic rename: several pytest-xdist workers may build at the same time."""
    deps = list(deps) + [ROOT / "oracle" / "liboracle.so", ROOT / "include" / "pairalign_b200.h"]
    if out.exists() and all(Path(d).stat().st_mtime <= out.stat().st_mtime for d in deps):
        return
    tmp = out.with_name(f"{out.name}.{os.getpid()}")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-o", str(tmp), *map(str, srcs),
           "-L", str(ROOT / "oracle"), "-loracle", "-Wl,-rpath," + str(ROOT / "oracle"),
           "-L", str(lib_dir), "-lpairalign_b200", "-Wl,-rpath," + str(lib_dir), *extra]
    subprocess.run(cmd, check=True, cwd=ROOT)
    os.replace(tmp, out)


@pytest.fixture(scope="module")
def exe():
    from phylommand_b200 import build
    from tests import oracle_lib
    build.build_library()
    oracle_lib.load()                                            # builds oracle/liboracle.so
    out = ROOT / "build" / "pairalign_hosttest"
    out.parent.mkdir(exist_ok=True)
    srcs = [HOST / "pairalign_main.cpp", HOST / "fasta_index.cpp", HOST / "mad_groups.cpp", HOST / "seqpair_batch.cpp",
            ROOT / "tests" / "host_double" / "seqpair_batch_oracle.cpp"]
    assert not any(s.name == "seqpair_batch_device.cpp" for s in srcs)
    _compile(out, srcs + sorted(HOST.glob("*.h")), srcs, build.LIB_DIR, ["-lpthread"])
    return out


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostcli")
    shutil.copytree(CLI_DIR / "inputs", d / "cli" / "inputs")
    shutil.copytree(ROOT / "tests" / "golden" / "example_files", d / "example_files")
    return d / "cli" / "inputs"


@pytest.mark.parametrize("entry", MANIFEST, ids=[e["tag"] for e in MANIFEST])
def test_host_side_matches_reference(exe, workdir, entry):
    cmd = [str(exe), *entry["flags"]] + ([entry["input"]] if entry["input"] else [])
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, timeout=900)
    assert r.returncode == entry["rc"], r.stderr.decode(errors="replace")[-2000:]
    want = (CLI_DIR / f"{entry['tag']}.out").read_bytes()
    if r.stdout != want:
        got_lines, want_lines = r.stdout.split(b"\n"), want.split(b"\n")
        for k, (g, w) in enumerate(zip(got_lines, want_lines)):
            if g != w:
                pytest.fail(f"line {k} differs:\n got  {g[:300]!r}\n want {w[:300]!r}\nstderr: {r.stderr.decode(errors='replace')[-500:]}")
        pytest.fail(f"length differs: got {len(r.stdout)} bytes, want {len(want)}")
    if entry.get("alignment_groups"):
        made = (workdir / (entry["input"] + ".alignment_groups")).resolve()
        assert made.read_bytes() == (CLI_DIR / f"{entry['tag']}.alignment_groups").read_bytes()
        made.unlink()


def test_bulk_rendering_of_a_large_batch(exe, tmp_path):
    """pairalign -a -n with more than a megabyte of text in one batch: the batch is rendered into one buffer by several
    host threads and written in one piece.  md5 of the unmodified reference's output for the same file
    (oracle/_ref/pairalign -a -n, 90 s on one core; synth.make_random(60, 77, 500, 700))."""
    import hashlib
    from phylommand_b200 import synth
    names, seqs = synth.make_random(60, 77, 500, 700)
    synth.write_fasta(tmp_path / "big_a.fst", names, seqs)
    r = subprocess.run([str(exe), "-a", "-n", "big_a.fst"], cwd=tmp_path, capture_output=True, timeout=900)
    assert r.returncode == 0 and len(r.stdout) == 2587172
    assert hashlib.md5(r.stdout).hexdigest() == "84fed2463faa907816700626d52cbd8f"


def test_product_binary_has_no_double():
    """The double lives under tests/ and the product build never sees it."""
    from phylommand_b200 import build
    import inspect
    assert "host_double" not in inspect.getsource(build)
    assert (HOST / "seqpair_batch_device.cpp").exists()


# ---- treeator -n: the host side of build/treeator_b200 (matrix reader, newick printer) ---------------------------------
NJ_GOLD = ROOT / "tests" / "golden" / "nj"
NJ_CASES = json.loads((NJ_GOLD / "manifest.json").read_text())


@pytest.fixture(scope="module")
def nj_exe():
    from phylommand_b200 import build
    from tests import oracle_lib
    build.build_library()
    oracle_lib.load()
    out = ROOT / "build" / "treeator_hosttest"
    out.parent.mkdir(exist_ok=True)
    srcs = [ROOT / "phylommand_b200" / "host_nj" / "treeator_nj_main.cpp", ROOT / "tests" / "host_double" / "nj_build_oracle.cpp"]
    _compile(out, srcs, srcs, build.LIB_DIR, [])
    return out


@pytest.mark.parametrize("case", NJ_CASES, ids=[c["tag"] for c in NJ_CASES])
def test_treeator_host_side_matches_reference(nj_exe, case):
    r = subprocess.run([str(nj_exe), *case["flags"], f"{case['tag']}.matrix"], cwd=NJ_GOLD, capture_output=True, timeout=600)
    assert r.returncode == case["rc"], r.stderr.decode(errors="replace")[-2000:]
    assert r.stdout == (NJ_GOLD / f"{case['tag']}.newick").read_bytes()


def test_treeator_host_side_stdin_and_ragged(nj_exe):
    data = (NJ_GOLD / "synth_ties_17.matrix").read_bytes()
    r = subprocess.run([str(nj_exe), "-n"], input=data, capture_output=True, timeout=600)
    assert r.returncode == 0 and r.stdout == (NJ_GOLD / "synth_ties_17.newick").read_bytes()
    r = subprocess.run([str(nj_exe), "-n"], input=b"a 1 2\nb 3 4\nc\n", capture_output=True, timeout=600)
    assert r.returncode == 1 and b"Error in distance matrix" in r.stderr
