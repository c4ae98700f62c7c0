"""The C-ABI library loads and exports every symbol include/gpucad_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from gpucadforam_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "gpucad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = declared()
    assert len(names) >= 45
    assert sorted(_capi.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in declared():
        assert hasattr(lib, n), n
    _capi.load()  # sets argtypes for all of them


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.check_output(["ldd", _capi.LIB_PATH]).decode()
    assert "liboracle" not in out and "gpucad_ref" not in out
    syms = subprocess.check_output(["nm", "-D", _capi.LIB_PATH]).decode()
    assert "orc_" not in syms


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.gcb_create(ctypes.byref(h), 0, None) != 0
    import gpucadforam_b200 as g
    import pytest
    with pytest.raises(RuntimeError):
        g.Context(0)
