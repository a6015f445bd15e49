"""The C-ABI shared library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from tensorlink_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tensorlink_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", src)))


def test_library_built():
    assert os.path.exists(native.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = native.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(native.exported_symbols()), set(syms) ^ set(native.exported_symbols())


def test_abi_version_and_error_string():
    lib = native.load()
    assert lib.tl_abi_version() == 1
    assert isinstance(native.last_error(), str)


def test_compute_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError):
        native.require_device()
    with pytest.raises(native.NativeError):
        native.rmsnorm_fwd(torch.zeros(1, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-6)
    sm, a, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert native.load().tl_device_info(sm, a, b) == -3          # TL_ERR_NO_DEVICE
    assert "no CUDA device" in native.last_error()
