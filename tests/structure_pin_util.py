"""Shared by tests/golden/make_reference_kind_structure.py (the generator) and tests/test_reference_pins.py: the
engine's side of one kind-structure pass as the compiled reference ran it (oracle/ref_kind_structure_main.cc) --
ephemeral kinds appended to an unassigned CrossCat, per round a cat-kernel sweep over all kinds with the debug taps on,
the proposer's score matrix, the block sampler, move_features and fresh ephemeral kinds; the ephemeral kinds dropped at
the end.  Test infrastructure."""
import numpy as np

from loom_b200 import Engine, sweep_actions
from loom_b200.engine import Model, RowBlock
import mixture_pin_util

STRIDE = mixture_pin_util.STRIDE


def build(case):
    one = dict(case, alpha=case["kind_py"][0], d=case["kind_py"][1])
    model1, table = mixture_pin_util.build(one)
    f2k = np.asarray(case["f2k"], np.uint32)
    K = int(f2k.max()) + 1
    model = Model(model1.features, f2k, kind_py=[tuple(case["kind_py"])] * K, topology=tuple(case["topology"]))
    return model, table


def round_actions(case, r):
    R = case["n_rows"]
    return sweep_actions(R) if r == 0 else sweep_actions(R, add=True, remove=True)


def run(abi, case, model, table):
    """-> dict(rounds=[dict(actions, picks [n][K], scores {(i, k): vector}, L [F][K], f2k, n_kinds, seatings [E][R])],
    final=dict(f2k, n_kinds, score_data, kinds=[(counts, ints, flts)]))"""
    E, seed = case["E"], case["seed"]
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, table))
    e.init_featureless_kinds(E, seed)                                   # KindKernel::KindKernel
    rounds = []
    for r in range(case["n_rounds"]):
        acts = round_actions(case, r)
        K = e.n_kinds()
        picks, scores = e.cat_sweep(acts, seed=seed + 10 + r, debug=True, debug_stride=STRIDE)   # add_row / remove_row
        vectors = []
        for i, a in enumerate(acts):
            if not a & 1:
                for k in range(K):
                    s = scores[i, k]
                    vectors.append(s[~np.isnan(s)])
        L = e.kind_proposer_score()                                     # try_run: infer_assignments' score phase
        e.kind_run(case["iterations"], seed + 100 + r)                  #          block sampler + move_features
        e.init_featureless_kinds(E, seed + 200 + r)                     #          init_featureless_kinds(empty_kind_count_)
        n_kinds, f2k = e.get_kinds()
        seatings = [e.get_assignments(k).astype(np.int64).tolist() for k in range(n_kinds - E, n_kinds)]
        rounds.append(dict(actions=acts, picks=picks, kinds_scored=K, vectors=vectors, L=L, f2k=f2k.tolist(), n_kinds=n_kinds,
                           seatings=seatings))
    e.init_featureless_kinds(0, seed)                                   # KindKernel::~KindKernel
    n_kinds, f2k = e.get_kinds()
    kinds = []
    for k in range(n_kinds):
        info = e.kind_info(k)
        counts, ints, flts = e.get_groups(k)
        kinds.append((counts[:info.n_groups].tolist(), ints[:, :info.n_groups], flts[:, :info.n_groups]))
    return dict(rounds=rounds, final=dict(f2k=f2k.tolist(), n_kinds=n_kinds, score_data=float(e.score_data()), kinds=kinds))


def compare(case, model, mine, ref, tol=2e-5):
    """the engine's run against the reference's printed one; returns the number of score vectors compared"""
    n = 0
    for r, (a, b) in enumerate(zip(mine["rounds"], ref["rounds"])):
        assert a["kinds_scored"] == b["kinds_scored"], (case["name"], r)
        assert len(a["vectors"]) == len(b["scores"]), (case["name"], r)
        for i, (x, y) in enumerate(zip(a["vectors"], b["scores"])):
            y = np.asarray(y, np.float64)
            assert len(x) == len(y), (case["name"], r, i)
            assert np.all(np.abs(x - y) <= tol * np.maximum(1.0, np.abs(y))), (case["name"], r, i)
            n += 1
        L = np.asarray(b["L"], np.float64)                              # [F][kinds]: one score vector per feature
        assert L.shape == a["L"].shape, (case["name"], r)
        assert np.all(np.abs(a["L"] - L) <= tol * np.maximum(1.0, np.abs(L))), (case["name"], r, float(np.abs(a["L"] - L).max()))
        assert a["f2k"] == b["f2k"] and a["n_kinds"] == b["n_kinds"], (case["name"], r)
    fa, fb = mine["final"], ref["final"]
    assert fa["f2k"] == fb["f2k"] and fa["n_kinds"] == fb["n_kinds"], case["name"]
    assert abs(fa["score_data"] - fb["score_data"]) <= tol * abs(fb["score_data"]), (case["name"], fa["score_data"], fb["score_data"])
    for k, ((counts, ints, flts), kb) in enumerate(zip(fa["kinds"], fb["kinds"])):
        assert counts == kb["counts"], (case["name"], k)
        feats = sorted(kb["features"], key=lambda f: f["feature"])
        assert [f["feature"] for f in feats] == model_kind_features(fa["f2k"], k)
        want_i = [row for f in feats for row in np.asarray(f["ints"], np.int64).reshape(len(counts), -1).T]
        want_f = [row for f in feats for row in np.asarray(f["flts"], np.float64).reshape(len(counts), -1).T]
        assert np.array_equal(np.asarray(want_i, np.int64).reshape(ints.shape), ints.astype(np.int64)), (case["name"], k)
        if len(want_f):
            assert np.allclose(np.asarray(want_f).reshape(flts.shape), flts, rtol=1e-9, atol=1e-9), (case["name"], k)
    return n


def model_kind_features(f2k, k):
    return [f for f, kk in enumerate(f2k) if kk == k]
