"""pairalign_b200: the all-pairs `pairalign` hot path of RybergGroup/phylommand on B200.

The product is the CUDA shared library behind include/pairalign_b200.h and the C++
command line built on it (phylommand_b200/host/).  `capi` binds the library for the
tests and bench.py; `synth` generates the synthetic inputs BASELINE.json names.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
