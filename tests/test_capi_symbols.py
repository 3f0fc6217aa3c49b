"""The C-ABI library loads and exports every symbol include/qpad_b200.h declares (no compute without a GPU)."""
import os
import re
import ctypes
import pytest
from qpad_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "qpad_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qpg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(capi.SIGNATURES) == _declared()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.QpadError, match="no CUDA device|CUDA"):
        capi.Ctx(64, 1, 0.1, 0.1)


def test_argument_errors():
    L = capi.load()
    h = ctypes.c_void_p()
    assert L.qpg_ctx_create(ctypes.byref(h), 0, None, 4, 1, 0.1, 0.1, capi.BND_OPEN, -1.0) == -1   # nr too small
    assert b"nr" in L.qpg_last_error()
    assert L.qpg_ctx_create(ctypes.byref(h), 0, None, 64, 9, 0.1, 0.1, capi.BND_OPEN, -1.0) == -1  # max_mode
    assert L.qpg_ctx_create(ctypes.byref(h), 0, None, 64, 1, 0.1, 0.1, 7, -1.0) == -1              # boundary
