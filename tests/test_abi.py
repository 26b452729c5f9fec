"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/waterlily_b200.h declares, reproduces the reference's constructor assertion, and refuses
to compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from waterlily_b200 import _cabi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "waterlily_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wl_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted("wl_" + k for k in _cabi.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(L, name), name


def test_oracle_exports_the_same_abi(oracle_lib):
    for name in header_symbols():
        assert hasattr(oracle_lib.lib, "wlo_" + name[3:]), name


def test_multilevel_size_assertion_message():
    # reference: test/maintests.jl:99
    with pytest.raises(_cabi.WLError, match="MultiLevelPoisson requires size=a2ⁿ, where n>2"):
        _cabi.Engine(lib.load(), (15, 81))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.WLError, match="no CUDA device"):
        _cabi.Engine(lib.load(), (16, 16))


def test_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "waterlily_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, f
