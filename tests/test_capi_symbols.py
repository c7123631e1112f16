"""CPU check that the C-ABI library loads and exports every symbol include/*.h declares."""

import os
import re

from qibojit_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qibojit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qj_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_capi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    for name in declared_symbols():
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.qj_version()


def test_no_cpu_fallback_without_device():
    import ctypes

    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _capi.load()
    out = ctypes.c_void_p()
    rc = lib.qj_create(0, None, ctypes.byref(out))
    assert rc == _capi.QJ_ERR_NODEVICE
    with pytest.raises(RuntimeError):
        _capi.check(rc)
    from qibojit_b200.backends.b200 import B200Backend

    with pytest.raises(RuntimeError):
        B200Backend()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qibojit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__never__", f"{f} mentions the oracle"
