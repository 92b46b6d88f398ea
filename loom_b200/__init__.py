"""loom_b200 -- B200-native CrossCat inference core behind loom's engine API.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of
include/loom_b200.h), `host/` (C++ shell re-hosting loom's classes), and this
Python mirror used by the tests and bench.py.
"""
from .engine import Engine, Model, RowBlock, sweep_actions, pack_mask, cat_sweep_multi, cat_sweep_streams, query_score  # noqa: F401
from ._abi import load_cuda, LoomError, UNASSIGNED  # noqa: F401
