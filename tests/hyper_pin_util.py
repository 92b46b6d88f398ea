"""Shared by tests/golden/make_reference_hyper.py (the generator) and tests/test_reference_pins.py: the engine's side of
a hyper-kernel run as the compiled reference ran it (oracle/ref_hyper_main.cc).  Test infrastructure."""
import numpy as np

from loom_b200 import Engine, sweep_actions
from loom_b200.engine import RowBlock
import structure_pin_util


def build(case):
    return structure_pin_util.build(case)


def run(abi, case, model, table):
    """-> (assignments [R][K] packed groups of the ADD-only sweep, topology (a, d), kinds [[a, d]], features [[hp0..3, aux...]])"""
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, table))
    picks, _ = e.cat_sweep(sweep_actions(case["n_rows"]), seed=case["seed"], debug=True)
    e.set_hyper_prior(case["grids"])
    e.hyper_run(case["seed"] + 7)
    specs, aux, kind_py, topo = e.get_hypers()
    feats = []
    for f in range(model.n_features):
        s = specs[f]
        n_aux = s.dim if s.type in (1, 2) else 0
        feats.append([float(s.hp[j]) for j in range(4)] + [float(a) for a in aux[s.aux_off:s.aux_off + n_aux]])
    return picks, [float(topo[0]), float(topo[1])], [[float(a), float(d)] for a, d in kind_py], feats


def compare(case, mine, ref):
    _, topo, kinds, feats = mine
    f32 = lambda xs: np.asarray(xs, np.float32)
    assert np.array_equal(f32(topo), f32(ref["topology"])), (case["name"], topo, ref["topology"])
    assert np.array_equal(f32(kinds), f32(ref["kinds"])), (case["name"], kinds, ref["kinds"])
    assert len(feats) == len(ref["features"])
    for f, (a, b) in enumerate(zip(feats, ref["features"])):
        assert np.array_equal(f32(a), f32(b)), (case["name"], f, a, b)
