"""B200-native batched Generals.io turn engine (drop-in for the reference's Engine.Step path).

The product is ``libgrlcuda.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/grlcuda.h``).  There is no CPU fallback: ``load_library()`` raises if the
CUDA library has not been built.
"""
from __future__ import annotations

import os

from . import _abi
from ._abi import BoundLibrary, Config, StepOutputs, StatePlanes, ACTION_DTYPE  # noqa: F401
from .engine import BatchedEngine, make_actions, make_config, set_action  # noqa: F401

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# GRL_LIB_PATH selects an experiment build of the same CUDA library (profiling variants)
LIB_PATH = os.environ.get("GRL_LIB_PATH") or os.path.join(_PKG_DIR, "csrc", "libgrlcuda.so")

_lib = None


def load_library() -> BoundLibrary:
    """Load the CUDA product library.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        _lib = BoundLibrary(LIB_PATH, "grl_")
    return _lib


def create_engine(**config) -> BatchedEngine:
    lib = load_library()
    return BatchedEngine(lib, make_config(lib, **config))
