"""The drop-in surface: build/pairalign_b200 must print byte-for-byte what the reference's
pairalign prints (tests/golden/cli/*.out, made by oracle/make_cli_golden.py from the unmodified
reference) for every output mode, matrix framing, clustering / alignment-group runs,
pair-fasta input, and the example files BASELINE.json config 0 names."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
CLI_DIR = ROOT / "tests" / "golden" / "cli"
MANIFEST = json.loads((CLI_DIR / "manifest.json").read_text())


@pytest.fixture(scope="module")
def exe():
    from phylommand_b200 import build
    build.build_library()
    path = build.build_cli()
    assert path is not None and path.exists()
    return path


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    shutil.copytree(CLI_DIR / "inputs", d / "cli" / "inputs")
    shutil.copytree(ROOT / "tests" / "golden" / "example_files", d / "example_files")
    return d / "cli" / "inputs"


@pytest.mark.parametrize("entry", MANIFEST, ids=[e["tag"] for e in MANIFEST])
def test_cli_matches_reference(exe, workdir, entry):
    cmd = [str(exe), *entry["flags"]] + ([entry["input"]] if entry["input"] else [])
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, timeout=600)
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


def test_cli_unknown_character_warnings(exe, workdir):
    """stderr carries the reference's per-pair 'Can not interpret' warnings (mixed.fst has U and ?)."""
    r = subprocess.run([str(exe), "-j", "-m", "mixed.fst"], cwd=workdir, capture_output=True, timeout=600)
    assert r.returncode == 0
    err = r.stderr.decode()
    # sequence 'delta' takes part in 10 pairs (11 distinct accessions); two unknown characters each time
    assert err.count("Can not interpret 'U'. Not in alphabet.") == 10
    assert err.count("Can not interpret '?'. Not in alphabet.") == 10


def test_cli_stdin_and_file_flag(exe, workdir):
    want = (CLI_DIR / "pure_j_n_m.out").read_bytes()
    data = (workdir / "pure.fst").read_bytes()
    r = subprocess.run([str(exe), "-j", "-n", "-m"], cwd=workdir, input=data, capture_output=True, timeout=600)
    assert r.stdout == want
    r = subprocess.run([str(exe), "-f", "pure.fst", "-j", "-n", "-m"], cwd=workdir, capture_output=True, timeout=600)
    assert r.stdout == want


def test_cli_argument_errors(exe, workdir):
    r = subprocess.run([str(exe), "--bogus", "pure.fst"], cwd=workdir, capture_output=True)
    assert r.returncode == 0 and b"Argument --bogus not recognized" in r.stderr and r.stdout == b""
    r = subprocess.run([str(exe), "--format", "xml", "pure.fst"], cwd=workdir, capture_output=True)
    assert r.returncode == 1
    r = subprocess.run([str(exe), "-g", "nonsense", "pure.fst"], cwd=workdir, capture_output=True)
    assert r.returncode == 1 and b"Do not recognize argument 'nonsense'" in r.stderr


def test_cli_two_devices_same_output(exe, workdir):
    """The in-process multi-device path (one host thread, stream set and pinned buffer per device entry).  On a box with
    one GPU the second entry is the same chip: the code path is the same."""
    import os
    import torch
    devices = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    env = dict(os.environ, PAIRALIGN_DEVICES=devices)
    for flags, golden in ((["-j", "-n", "-m"], "pure_j_n_m.out"), (["-a", "-n"], "pure_a_n.out")):
        want = (CLI_DIR / golden).read_bytes()
        r = subprocess.run([str(exe), *flags, "pure.fst"], cwd=workdir, capture_output=True, timeout=600, env=env)
        assert r.returncode == 0 and r.stdout == want, (flags, r.stderr[-500:])
