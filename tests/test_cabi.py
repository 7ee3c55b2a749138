"""The C-ABI library loads on a CPU-only box and exports every symbol include/rfk.h declares.
No compute call is made here (that needs a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rfk.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rfk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import recurrent_flows_msc_b200 as r
    if not os.path.exists(r._lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return r._lib


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "rfk_conv_gemm" in syms and "rfk_squeeze2d" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(lib):
    cdll = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, f"librfk.so lacks {missing}"


def test_binding_covers_header(lib):
    bound = set(lib.SIGNATURES) | {"rfk_last_error"}
    assert bound == set(declared_symbols())


def test_version_and_error_string(lib):
    l = lib.lib()
    assert l.rfk_version() == 1
    assert isinstance(l.rfk_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call, with a message."""
    l = lib.lib()
    rc = l.rfk_squeeze2d(None, None, 1, 1, 2, 2, 0, None)
    assert rc == -1 and b"rfk_squeeze2d" in l.rfk_last_error()
    rc = l.rfk_conv_gemm(None, 1, 2, 2, 64, 64, None, 16, 16, 9, None, None, 0, 0, None, 64, 0, None)
    assert rc == -1
    with pytest.raises(lib.RfkError):
        lib.call("rfk_coupling_tail", None, None, 1, 3, 4, 0, None, None, None, 0, None)


def test_product_path_refuses_cpu_tensors():
    import torch
    import recurrent_flows_msc_b200 as r
    with torch.no_grad(), pytest.raises(RuntimeError):
        r.Squeeze2d()(torch.zeros(1, 1, 2, 2), undo_squeeze=False)
    with torch.no_grad(), pytest.raises(RuntimeError):
        r.ConvLSTM(2, 2, [3, 3])(torch.zeros(1, 1, 2, 4, 4))
