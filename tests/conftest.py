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


# The GPU tests of the last file, surest first: under `pytest -x` a failure in a late, never-run-before test must not
# keep the earlier ones from reporting.
_ZZ_ORDER = ["test_query_sample_matches_oracle_on_the_gpu", "test_loom_query_on_the_gpu", "test_score_error_per_feature_type",
             "test_query_client_on_the_gpu", "test_masked_hyper_run_on_one_gpu",
             "test_fused_query_score_matches_the_two_kernel_path", "test_differ_on_the_gpu",
             "test_cuda_reproduces_query_and_differ_fixture", "test_loom_tare_and_loom_sparsify_on_the_gpu",
             "test_generate_rows_matches_oracle_on_the_gpu", "test_loom_generate_and_loom_mix_on_the_gpu",
             "test_row_sharded_accumulate_nccl",
             "test_feature_sharded_hyper_nccl"]


def pytest_collection_modifyitems(config, items):
    def key(indexed):
        i, item = indexed
        if item.fspath.basename == "test_zz_query_gpu.py":
            name = item.originalname if hasattr(item, "originalname") and item.originalname else item.name
            return (1, _ZZ_ORDER.index(name) if name in _ZZ_ORDER else len(_ZZ_ORDER), i)
        return (0, 0, i)
    items[:] = [item for _, item in sorted(enumerate(items), key=key)]
