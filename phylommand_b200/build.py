"""Build the native pieces in-tree (no JIT cache): the sm_100a CUDA module behind
include/pairalign_b200.h and, for tests only, the oracle under oracle/.

    python -m phylommand_b200.build          # everything
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "phylommand_b200"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libpairalign_b200.so"
CLI_PATH = ROOT / "build" / "pairalign_b200"
NJ_CLI_PATH = ROOT / "build" / "treeator_b200"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA module cannot be built")
    return exe


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -> phylommand_b200/lib/libpairalign_b200.so (sm_100a only)."""
    srcs = sorted((PKG / "csrc").glob("*.cu"))
    deps = srcs + sorted((PKG / "csrc").glob("*.cuh")) + [ROOT / "include" / "pairalign_b200.h"]
    if not force and _newer(LIB_PATH, deps):
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", str(LIB_PATH), *map(str, srcs), "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=ROOT)
    return LIB_PATH


def build_cli(force: bool = False) -> Path | None:
    """The C++ host driver (pairalign command line) linked against the library."""
    host = PKG / "host"
    srcs = sorted(host.glob("*.cpp"))
    if not srcs:
        return None
    deps = srcs + sorted(host.glob("*.h")) + [ROOT / "include" / "pairalign_b200.h", LIB_PATH]
    if not force and _newer(CLI_PATH, deps):
        return CLI_PATH
    CLI_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-o", str(CLI_PATH),
           *map(str, srcs), "-L", str(LIB_DIR), "-lpairalign_b200",
           "-Wl,-rpath," + str(LIB_DIR), "-Wl,-rpath,$ORIGIN/../phylommand_b200/lib", "-lpthread"]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return CLI_PATH


def build_nj_cli(force: bool = False) -> Path:
    """`treeator -n` host program (neighbour joining of the -m matrix) linked against the library."""
    srcs = sorted((PKG / "host_nj").glob("*.cpp"))
    deps = srcs + [ROOT / "include" / "pairalign_b200.h", LIB_PATH]
    if not force and _newer(NJ_CLI_PATH, deps):
        return NJ_CLI_PATH
    NJ_CLI_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-o", str(NJ_CLI_PATH),
           *map(str, srcs), "-L", str(LIB_DIR), "-lpairalign_b200",
           "-Wl,-rpath," + str(LIB_DIR), "-Wl,-rpath,$ORIGIN/../phylommand_b200/lib"]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return NJ_CLI_PATH


def build_oracle() -> None:
    """Test infrastructure: the C restatement and (when /root/reference exists) the
    unmodified reference compiled into oracle/_ref/.  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "all"], check=True)


def build_all(force: bool = False) -> None:
    build_library(force=force)
    build_cli(force=force)
    build_nj_cli(force=force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB_PATH)
