"""The measurement plumbing that can be checked without a device: bench.py's auxiliary experiments run in processes of
their own and report errors instead of raising; the library override used by the build-variant A/B is loud when wrong;
the reference arm prints the contract's JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench(oracle_abi):
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(REPO, "bench.py"))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
    finally:
        sys.argv = argv
    return module


def test_auxiliary_experiments_are_isolated_and_budgeted(bench, monkeypatch):
    ok = bench.measure_in_subprocess("differ_bench.py", [2000, "oracle"])
    assert ok["impl"] == "oracle" and ok["tare"]["ms"] > 0 and ok["sparsify"]["rows_per_s"] > 0
    failed = bench.measure_in_subprocess("query_bench.py", [100])            # too few arguments: the script dies
    assert "error" in failed and "exit" in failed["error"]
    slow = bench.measure_in_subprocess("differ_bench.py", [200000, "oracle"], timeout=1)
    assert "timed out" in slow["error"]
    monkeypatch.setattr(bench, "AUX_BUDGET_S", 0.0)
    assert "skipped" in bench.measure_in_subprocess("differ_bench.py", [2000, "oracle"])


def test_library_override_is_loud(tmp_path):
    code = ("import sys; sys.path.insert(0, %r); from loom_b200 import _abi\n"
            "try:\n    _abi.load_cuda()\nexcept _abi.LoomError as e:\n    print('LOUD', e)\n" % REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         env=dict(os.environ, LOOM_B200_LIB=str(tmp_path / "missing.so")))
    assert "LOUD" in out.stdout and "does not exist" in out.stdout


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--rows", "4000", "--ref-rows", "1500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cells/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "C3" and line["config"]["ephemeral_kinds"] == 32
