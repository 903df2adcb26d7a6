"""integration/pairalign_b200.patch: the reference's OWN pairalign.cpp wired to include/pairalign_b200.h.

Without a GPU (this file's unmarked tests, run where /root/reference exists):
  * the patch applies cleanly to a scratch copy of the reference's src/ and leaves the default build untouched
    (`-UPAIRALIGN_B200`: same preprocessed program);
  * with -DPAIRALIGN_B200 every pairalign source passes `g++ -fsyntax-only` against the header and the glue
    (integration/b200_batch.h), and links against the in-tree library;
  * that binary has no CPU path: without a device it exits with status 2 and says so;
  * linked in front of a test double of the compute entry points (tests/host_double/capi_oracle_double.cpp, answers
    from the oracle) it prints byte for byte what the unmodified reference printed for the FASTA command lines of
    tests/golden/cli/ -- i.e. the collect / align / replay restructuring of cluster() is observably the same loop.
On the GPU box (-m gpu): the same binary linked against the real module (oracle/_ref/pairalign_b200_patched, built by
oracle/Makefile here, travels like the other reference builds) against the same goldens."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference/src")
PATCH = ROOT / "integration" / "pairalign_b200.patch"
CLI_DIR = ROOT / "tests" / "golden" / "cli"
MANIFEST = [e for e in json.loads((CLI_DIR / "manifest.json").read_text())
            if e["input"] and e["input"].endswith(".fst") and "--format" not in e["flags"]]
SRCS = ["seqpair.cpp", "pairalign.cpp", "align_group.cpp", "seqdatabase.cpp", "argv_parser.cpp", "indexedfasta.cpp"]
PATCHED = ROOT / "oracle" / "_ref" / "pairalign_b200_patched"

needs_reference = pytest.mark.skipif(not REF_SRC.exists(), reason="the reference checkout is not on this machine")


@pytest.fixture(scope="module")
def patched_src(tmp_path_factory):
    d = tmp_path_factory.mktemp("refpatch")
    shutil.copytree(REF_SRC, d / "src")
    r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", str(PATCH)], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout
    return d / "src"


FLAGS = ["-std=c++11", "-w", "-DPAIRALIGN_B200", "-I", str(ROOT / "include"), "-I", str(ROOT / "integration")]


@needs_reference
def test_patch_applies_and_compiles_against_the_header(patched_src):
    for src in SRCS:
        r = subprocess.run(["g++", *FLAGS, "-fsyntax-only", src], cwd=patched_src, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    # without the macro the patched file is the reference's program: same preprocessed text
    pre = [subprocess.run(["g++", "-std=c++11", "-w", "-E", "-P", "pairalign.cpp"], cwd=c, capture_output=True, text=True).stdout
           for c in (patched_src, REF_SRC)]
    assert pre[0] == pre[1] and len(pre[0]) > 10000
    assert "pairalign_b200:" in (patched_src / "Makefile").read_text()


@needs_reference
def test_patched_reference_has_no_cpu_path(patched_src, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phylommand_b200 import build
    build.build_library()
    exe = tmp_path / "pairalign_b200_real"
    subprocess.run(["g++", *FLAGS, "-O1", "-o", str(exe), *SRCS, "-L", str(build.LIB_DIR), "-lpairalign_b200",
                    "-Wl,-rpath," + str(build.LIB_DIR)], cwd=patched_src, check=True)
    r = subprocess.run([str(exe), "-j", "-n", "-m", "pure.fst"], cwd=CLI_DIR / "inputs", capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr and r.stdout == ""


@pytest.fixture(scope="module")
def double_exe(patched_src, tmp_path_factory):
    from phylommand_b200 import build
    from tests import oracle_lib
    build.build_library()
    oracle_lib.load()
    exe = tmp_path_factory.mktemp("refbin") / "pairalign_b200_double"
    subprocess.run(["g++", *FLAGS, "-O2", "-o", str(exe), *SRCS, str(ROOT / "tests" / "host_double" / "capi_oracle_double.cpp"),
                    "-L", str(ROOT / "oracle"), "-loracle", "-Wl,-rpath," + str(ROOT / "oracle"),
                    "-L", str(build.LIB_DIR), "-lpairalign_b200", "-Wl,-rpath," + str(build.LIB_DIR)], cwd=patched_src, check=True)
    return exe


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("patchcli")
    shutil.copytree(CLI_DIR / "inputs", d / "cli" / "inputs")
    shutil.copytree(ROOT / "tests" / "golden" / "example_files", d / "example_files")
    return d / "cli" / "inputs"


def _check(exe, workdir, entry):
    r = subprocess.run([str(exe), *entry["flags"], entry["input"]], cwd=workdir, capture_output=True, timeout=900)
    assert r.returncode == entry["rc"], r.stderr.decode(errors="replace")[-2000:]
    assert r.stdout == (CLI_DIR / f"{entry['tag']}.out").read_bytes(), entry["tag"]
    if entry.get("alignment_groups"):
        made = (workdir / (entry["input"] + ".alignment_groups")).resolve()
        assert made.read_bytes() == (CLI_DIR / f"{entry['tag']}.alignment_groups").read_bytes()
        made.unlink()


@needs_reference
@pytest.mark.parametrize("entry", [e for e in MANIFEST if not e["tag"].startswith("example_")], ids=lambda e: e["tag"])
def test_patched_reference_on_the_double_matches_the_reference(double_exe, workdir, entry):
    _check(double_exe, workdir, entry)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", MANIFEST, ids=lambda e: e["tag"])
def test_patched_reference_on_the_cuda_module_matches_the_reference(workdir, entry):
    if not PATCHED.exists():
        pytest.skip("oracle/_ref/pairalign_b200_patched did not travel (built by oracle/Makefile where the reference exists)")
    from phylommand_b200 import build
    build.build_library()
    _check(PATCHED, workdir, entry)
