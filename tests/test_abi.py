"""CPU suite: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/rala_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from rala_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rala_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rala_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(api.EXPORTS) == set(names), set(api.EXPORTS) ^ set(names)
    assert lib.rala_b200_abi_version() == 1


def test_no_device_means_failure_not_fallback():
    """Without a CUDA device the context constructor fails loudly; nothing routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        return
    lib = ctypes.CDLL(build.build())
    handle = ctypes.c_void_p()
    assert lib.rala_b200_create(ctypes.byref(handle), 0) != 0 and not handle.value
    try:
        api.Context(0)
    except api.RalaB200Error:
        pass
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rala_b200/ may import, include or load it."""
    pkg = os.path.join(ROOT, "rala_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s*[\"<].*oracle|liboracle|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f
