"""CPU-side checks of the drop-in boundary: the CUDA library loads and exports
every symbol include/loom_b200.h declares (no compute without a GPU), and the
oracle exports the same ABI under the lo_ prefix."""
import os
import re

import numpy as np
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


def test_comm_header_and_binding_agree_and_are_exported():
    """include/loom_b200_comm.h: every NCCL entry point is bound (CUDA library only) and exported"""
    text = open(os.path.join(REPO, "include", "loom_b200_comm.h")).read()
    declared = sorted(set(re.findall(r"\bint (lb_comm_\w+)\(", text)))
    assert declared == sorted("lb_" + n for n in _abi.COMM_SIGNATURES)
    abi = _abi.load_cuda()
    for name in declared:
        assert hasattr(abi.lib, name), name


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


def test_cpp_host_shell_compiles_and_links(tmp_path):
    """The C++ shell with loom's class names (loom_b200/host/loom_b200.hpp) builds against the library."""
    import subprocess
    test_cuda_library_exports_every_symbol()
    exe = tmp_path / "host_shell_example"
    csrc = os.path.join(REPO, "loom_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", os.path.join(REPO, "tests", "host_shell_example.cc"),
                           "-o", str(exe), "-L" + csrc, "-lloomb200", "-Wl,-rpath," + csrc,
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    subprocess.check_call([str(exe)])   # without "run": link check only


@pytest.mark.gpu
def test_cpp_host_shell_runs_on_the_gpu(tmp_path):
    """CrossCat / CatKernel / MultiCatKernel / the schedule driver of the C++ shell, executed."""
    import subprocess
    exe = tmp_path / "host_shell_example"
    csrc = os.path.join(REPO, "loom_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", os.path.join(REPO, "tests", "host_shell_example.cc"),
                           "-o", str(exe), "-L" + csrc, "-lloomb200", "-Wl,-rpath," + csrc,
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.check_output([str(exe), "run"], timeout=120).decode()
    lines = out.strip().splitlines()
    assert lines[0].startswith("score ") and lines[1].startswith("scores ") and lines[2].startswith("batches ")
    assert int(lines[2].split()[1]) >= 3
    assert all(np.isfinite(float(x)) for x in lines[1].split()[1:])
