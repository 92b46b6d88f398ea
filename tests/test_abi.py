"""CPU-side checks of the drop-in boundary: the CUDA library loads and exports
every symbol include/loom_b200.h declares (no compute without a GPU), and the
oracle exports the same ABI under the lo_ prefix."""
import os
import re

import pytest

from loom_b200 import _abi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "loom_b200.h")).read()
    return sorted(set(re.findall(r"LB_NAME\((\w+)\)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_abi.SIGNATURES)


def test_cuda_library_exports_every_symbol():
    if not os.path.exists(_abi.CUDA_LIB):
        import __graft_entry__
        __graft_entry__.build()
    abi = _abi.load_cuda()
    for name in declared_symbols():
        assert hasattr(abi.lib, "lb_" + name), name


def test_oracle_exports_same_abi(oracle_abi):
    for name in declared_symbols():
        assert hasattr(oracle_abi.lib, "lo_" + name), name


def test_no_device_is_an_error_not_a_fallback():
    import ctypes
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    abi = _abi.load_cuda()
    h = ctypes.c_void_p()
    assert abi.create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in abi.last_error()
