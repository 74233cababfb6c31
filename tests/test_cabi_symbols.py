"""CPU: the C-ABI library builds, loads and exports every symbol include/neat_b200.h declares."""
import os
import re

from neat_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "neat_b200.h")).read()
    declared = set(re.findall(r"\b(neat_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_create_fails_loudly_without_gpu():
    import ctypes
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    cfg = _lib.NetConfig(9, 256, 4, 6, 256, 5, 256, 4, 3.0, 20.0)
    h = ctypes.c_void_p()
    assert lib.neat_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert lib.neat_last_error()
