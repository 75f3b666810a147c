"""ragnar_b200 — B200-native (sm_100a) implementation of haykh/ragnar's radiation
hot path behind ragnar's unchanged Python API.

    import ragnar_b200
    rg = ragnar_b200.load()      # the compiled pybind11 module `ragnar`
    rg.Initialize()

``import ragnar`` works directly once this package directory is on ``sys.path``
(``ragnar_b200.load()`` puts it there).  The lower-level C-ABI of
``libragnar_cuda.so`` (include/ragnar_cuda.h) is bound with ctypes in
``ragnar_b200.cabi``; ``ragnar_b200.dist`` holds the one-process-per-GPU plumbing.

There is no CPU fallback: without the compiled extension ``load()`` raises, and
without a CUDA device ``rg.Initialize()`` raises.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent

__all__ = ["load", "PKG_DIR"]


def load():
    """Import and return the compiled ``ragnar`` extension module (built in-tree
    by ``python -m ragnar_b200.build``)."""
    if str(PKG_DIR) not in sys.path:
        sys.path.insert(0, str(PKG_DIR))
    try:
        return importlib.import_module("ragnar")
    except ImportError as exc:  # fail loudly: never substitute a Python path
        raise ImportError(
            "the compiled `ragnar` extension is missing or broken — run "
            "`python -m ragnar_b200.build` (needs nvcc); there is no Python/CPU fallback"
        ) from exc
