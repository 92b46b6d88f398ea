"""Shared by tests/golden/make_reference_cat_kernel.py (the generator) and tests/test_reference_pins.py: the engine's
side of a multi-kind cat-kernel sweep as the compiled reference CatKernel ran it (oracle/ref_cat_kernel_main.cc).
Test infrastructure."""
import numpy as np

from loom_b200 import Engine
from loom_b200.engine import RowBlock
import mixture_pin_util
import structure_pin_util

STRIDE = mixture_pin_util.STRIDE
UNASSIGNED = 0xFFFFFFFF


def build(case):
    model, table = structure_pin_util.build(case)
    if case.get("tare"):                  # booleans that are nearly constant, so that a tare row exists
        rng = np.random.RandomState(case["seed"] + 1)
        for row in table:
            for f in range(min(4, model.n8())):
                if model.features[f]["type"] == "bb" and rng.uniform() < 0.9:
                    row[f] = 1 if f == 1 else 0
    return model, table


def run(abi, case, model, table, tare=None):
    """-> (actions, picks [n][K], score vectors of every (ADD, kind) in order, score_data, per kind (counts, packed group per row))"""
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, table))
    if tare is not None:
        obs, val = tare
        d = e.sparsify(obs, val)
        d["row_has_tare"] = np.ones(len(table), np.uint8)
        e.set_rows_diff(**d)
    acts = mixture_pin_util.actions_of(case)
    picks, scores = e.cat_sweep(acts, seed=case["seed"], debug=True, debug_stride=STRIDE)
    K = model.n_kinds
    vectors = []
    for i, a in enumerate(acts):
        if not a & 1:
            for k in range(K):
                s = scores[i, k]
                vectors.append(s[~np.isnan(s)])
    kinds = []
    for k in range(K):
        info = e.kind_info(k)
        counts, _, _ = e.get_groups(k)
        kinds.append((counts[:info.n_groups].tolist(), e.get_assignments(k)))
    return acts, picks, vectors, float(e.score_data()), kinds


def compare(case, mine, ref, tol=2e-5):
    acts, picks, vectors, score_data, kinds = mine
    assert len(vectors) == len(ref["scores"]), case["name"]
    for i, (x, y) in enumerate(zip(vectors, ref["scores"])):
        y = np.asarray(y, np.float64)
        assert len(x) == len(y), (case["name"], i)
        assert np.all(np.abs(x - y) <= tol * np.maximum(1.0, np.abs(y))), (case["name"], i)
    assert abs(score_data - ref["score_data"]) <= tol * abs(ref["score_data"]), (case["name"], score_data, ref["score_data"])
    rows = [rid - 1000 for rid in ref["rowids"]]                    # the rows still assigned, in FIFO order
    for k, ((counts, assign), kb) in enumerate(zip(kinds, ref["kinds"])):
        assert counts == kb["counts"], (case["name"], k)
        assert sorted(rows) == np.flatnonzero(assign != UNASSIGNED).tolist(), (case["name"], k)
        assert assign[rows].tolist() == kb["packed"], (case["name"], k)
        # the queue holds GLOBAL ids: stable names, one per group that ever existed (src/cat_kernel.hpp:222)
        assert len(set(kb["global"])) == len(set(kb["packed"]))
    return len(vectors)
