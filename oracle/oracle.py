"""Python handle on the CPU oracle (oracle/wl_oracle.cpp) — TEST INFRASTRUCTURE ONLY.

Importers: tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs.  It binds ``liboracle.so`` through the same ctypes table as the product (prefix ``wlo_``) and
re-uses the package's host logic, so a parity test drives both sides with identical calls.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from waterlily_b200._cabi import CABI, Engine
from waterlily_b200.sim import _SimulationBase

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "wl_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)


def load() -> CABI:
    global _lib
    if _lib is None:
        build()
        _lib = CABI(LIB_PATH, "wlo_")
        _lib.lib.wlo_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(load().lib.wlo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle from here on (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    load().lib.wlo_set_num_threads(int(n))
    return num_threads()


def engine(dims, **kw) -> Engine:
    return Engine(load(), dims, **kw)


class OracleSimulation(_SimulationBase):
    """The reference's ``Simulation`` with ``mem=Array`` semantics, computed by the oracle."""

    @classmethod
    def _lib(cls):
        return load()


def debug_flux(which: str, f, I: int = 0, Ip: int = 0, u: float = 0.0) -> float:
    """The oracle's flux helpers on a 1-D vector, 1-based indices like the reference (`quick`, `phi`, `phiu`, `phiuP`,
    `phiuL`, `phiuR` = quick, ϕ, ϕu, ϕuP, ϕuL, ϕuR of src/Flow.jl:3-9) — for the pins of maintests.jl:120-155."""
    import ctypes
    import numpy as np
    fn = load().lib.wlo_debug_flux
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.c_double]
    a = np.ascontiguousarray(f, dtype=np.float32)
    k = ["quick", "phi", "phiu", "phiuP", "phiuL", "phiuR"].index(which)
    return float(fn(k, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(I) - 1, int(Ip) - 1, float(u)))
