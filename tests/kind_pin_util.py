"""Shared by tests/golden/make_reference_kind_kernel.py (the generator) and tests/test_reference_pins.py: the oracle's
side of the KindKernel life cycle the compiled reference was driven through (oracle/ref_kind_kernel_main.cc) --
construction, try_run on supplied score matrices, destruction.  Test infrastructure."""
import ctypes

import numpy as np

from loom_b200 import Engine, synth, sweep_actions
from loom_b200.engine import Model


def build_engine(abi, case):
    """a small assigned engine with the case's initial feature -> kind map (the data only has to exist: the scores the
    kind kernel sees are the case's, not the data's)"""
    feats = synth.make_features([tuple(t) for t in case["types"]])
    f2k = np.asarray(case["f2k"], np.uint32)
    K = int(f2k.max()) + 1
    topology = (case["alpha"], case["d"])
    shape = Model(feats, f2k, kind_py=[synth.CLUSTERING] * K, topology=topology)
    _, rows, _ = synth.generate(case["n_rows"], [tuple(t) for t in case["types"]], density=0.7, seed=case["seed"], f2k=f2k)
    e = Engine(abi)
    e.set_model(shape)
    e.set_rows(rows)
    e.cat_sweep(sweep_actions(case["n_rows"]), seed=case["seed"])
    return e


def set_scores(abi, e, by_position):
    """by_position [n_kinds][F] (row p = the kind at position p, as the reference harness takes them)"""
    lib = abi.lib
    lib.lo_set_proposer_scores.restype = ctypes.c_int
    lib.lo_set_proposer_scores.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L = np.ascontiguousarray(np.asarray(by_position, np.float32).T)          # [F][K]
    assert L.shape == (e.model.n_features, e.n_kinds()), (L.shape, e.n_kinds())
    abi.check(lib.lo_set_proposer_scores(e.h, L.ctypes.data))


def life_cycle(abi, case, scores_for_round):
    """-> the states [(step, f2k, n_kinds)] after construction, every try_run and destruction.  scores_for_round(r,
    n_kinds) supplies round r's matrix (the generator draws it, the test reads the fixture)."""
    e = build_engine(abi, case)
    E, seed = case["E"], case["seed"]
    e.init_featureless_kinds(E, seed)                       # KindKernel::KindKernel (src/kind_kernel.cc:66-69)
    states = [("constructed",) + _state(e)]
    for r in range(case["n_rounds"]):
        e.kind_proposer_score()
        set_scores(abi, e, scores_for_round(r, e.n_kinds()))
        e.kind_run(case["iterations"], seed + 100 + r)      # try_run: infer_assignments + move_features
        e.init_featureless_kinds(E, seed + 200 + r)         #          + init_featureless_kinds(empty_kind_count_)
        states.append(("try_run",) + _state(e))
    e.init_featureless_kinds(0, seed)                       # KindKernel::~KindKernel (src/kind_kernel.cc:74-80)
    states.append(("destroyed",) + _state(e))
    return states


def _state(e):
    n, f2k = e.get_kinds()
    return f2k.tolist(), n
