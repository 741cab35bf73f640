"""the C-ABI library builds, loads, and exports every symbol include/mpb200.h declares (no compute without a GPU)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_symbols():
    from multiprime_b200 import _lib, build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    header = open(os.path.join(ROOT, "include", "mpb200.h")).read()
    declared = set(re.findall(r"\b(mpb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().mpb_abi_version() == 3


def test_no_cpu_fallback():
    """without a device the product path must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from multiprime_b200 import _lib
    with pytest.raises(_lib.MpbError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multiprime_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
