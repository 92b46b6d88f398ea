import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long statistical test")


@pytest.fixture(scope="session")
def oracle_abi():
    """CPU oracle bound through the same ctypes class as the product."""
    from loom_b200 import _abi
    path = os.path.join(REPO, "oracle", "libloom_oracle.so")
    srcs = [os.path.join(REPO, "oracle", f) for f in os.listdir(os.path.join(REPO, "oracle"))
            if f.endswith((".c", ".h"))] + [os.path.join(REPO, "include", "loom_b200.h")]
    if (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
    return _abi.Abi(path, "lo_")


@pytest.fixture(scope="session")
def cuda_abi():
    import ctypes
    from loom_b200 import _abi
    abi = _abi.load_cuda()  # raises if the .so is missing: no CPU fallback
    return abi


def has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int()
        return lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False
