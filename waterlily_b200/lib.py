"""Loader of ``libwaterlily_b200.so`` — the only compute backend of this package.

There is deliberately no CPU fallback: if the CUDA library is missing or cannot be loaded the
import fails loudly (build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C waterlily_b200/csrc``).
"""
from __future__ import annotations

import os

from ._cabi import CABI

# WL_LIB_PATH: a differently-built copy of the same library (kernel tuning experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("WL_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libwaterlily_b200.so")
_lib = None


def load() -> CABI:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                "(run __graft_entry__.build() or `make -C waterlily_b200/csrc`). "
                "waterlily_b200 has no CPU fallback.")
        _lib = CABI(LIB_PATH, "wl_")
    return _lib
