"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/sofa_b200.h declares, and refuses to run without a device (no CPU path)."""
import ctypes as C
import os
import re

import pytest

import sofa_b200
from sofa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sofa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sofab200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = sofa_b200.load()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sofa_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "python binding table and header disagree"


def test_version_and_error_string():
    L = sofa_b200.load()
    assert b"sm_100a" in L.sofab200_version()
    assert isinstance(L.sofab200_last_error(), bytes)


def test_no_cpu_path():
    """Without a CUDA device every entry point fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = sofa_b200.load()
    h = C.c_void_p()
    rc = L.sofab200_ctx_create(0, None, C.byref(h))
    assert rc == -3 and b"no CPU path" in L.sofab200_last_error()
    with pytest.raises(_lib.Sofab200Error):
        sofa_b200.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under sofa_b200/ may import, include or link oracle/ (or tests/)."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "sofa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle_lib|sofa_oracle|oracle/|emu_lib", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
