"""Tare-compressed input (ProductValue::Diff): the worked example of doc/adapting.md:323-414 and
equivalence of the diff path with the dense path (score_diff == score_value of the
reconstructed row; add_diff == add_value)."""
import numpy as np
import pytest

from loom_b200 import Engine, Model, RowBlock, sweep_actions
from loom_b200 import synth, tare


def bool_model(n):
    return Model([dict(type="bb", alpha=0.5, beta=0.5)] * n, [0] * n)


def test_adapting_md_worked_example():
    # doc/adapting.md:323-414: row [F, T, -, F, F]; tare observed [1,1,0,0,0] values [F, F]
    model = bool_model(5)
    table = [[0, 1, None, 0, 0]]
    rows = RowBlock.from_table(model, table)
    tare_obs = np.uint8([1, 1, 0, 0, 0])
    tare_val = np.uint32([0, 0, 0, 0, 0])
    d = tare.sparsify(model, rows, tare_obs, tare_val)
    assert list(d["pos_feature"]) == [1, 3, 4] and list(d["pos_value"]) == [1, 0, 0]   # pos sparse [1,3,4] = [T,F,F]
    assert list(d["neg_feature"]) == [1]                                                # neg sparse [1] = [F]
    assert list(d["pos_ptr"]) == [0, 3] and list(d["neg_ptr"]) == [0, 1]


def test_make_tare_rules():
    # mode must cover > 50 % of ALL rows; reals never tared; counts only with mode < 16
    model, rows, _ = synth.generate(400, [("bb", 2), ("gp", 2), ("nich", 1)], density=1.0, seed=1, single_kind=True)
    rows.cat8[0, :] = 0; rows.cat8[0, :150] = 1            # 62.5 % zeros -> tared to 0
    rows.cat8[1, :] = 0; rows.cat8[1, :200] = 1            # exactly 50 % -> not tared
    rows.wide[0, :] = 3; rows.wide[0, :100] = 7            # gp mode 3 (75 %) -> tared
    rows.wide[1, :] = 40                                   # gp mode >= 16 -> never seen by the summary
    to, tv = tare.make_tare(model, rows)
    assert list(to) == [1, 0, 1, 0, 0] and tv[0] == 0 and tv[2] == 3


def _equivalence(abi):
    model, rows, _ = synth.generate(600, [("bb", 6), ("dd4", 2), ("gp", 2), ("nich", 2)], density=0.8, seed=4)
    rows.cat8[:4] = (np.random.default_rng(0).random((4, 600)) < 0.05)   # sparse booleans: tare = 0
    rows.wide[0, :] = 3; rows.wide[0, ::7] = 25; rows.wide[0, ::11] = 5   # a tared count column (mode 3) with counts beyond 16
    to, tv = tare.make_tare(model, rows)
    assert to[:4].all() and to[model.n8()] == 1 and tv[model.n8()] == 3
    d = tare.sparsify(model, rows, to, tv)
    dense, diff = Engine(abi), Engine(abi)
    for e in (dense, diff):
        e.set_model(model)
    dense.set_rows(rows)
    diff.set_rows_diff(**d)
    acts = np.concatenate([sweep_actions(600), sweep_actions(600, add=True, remove=True)])
    for e in (dense, diff):
        e.cat_sweep(acts, seed=8)
    for k in range(model.n_kinds):
        a, b = dense.get_groups(k), diff.get_groups(k)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(dense.get_assignments(k), diff.get_assignments(k))
        assert np.array_equal(dense.score_rows(k), diff.score_rows(k))
        # score_diff proper: tare score + pos - neg from the compressed form == the decoded row's score_value
        sparse, full = diff.score_rows_diff(k).astype(np.float64), diff.score_rows(k).astype(np.float64)
        assert sparse.shape == full.shape
        assert (np.abs(sparse - full) <= 1e-5 * np.maximum(1.0, np.abs(full))).all(), (k, np.abs(sparse - full).max())
        part = diff.score_rows_diff(k, 37, 215)
        assert np.array_equal(part, diff.score_rows_diff(k)[37:215])
    assert abs(dense.score_data() - diff.score_data()) <= 1e-12 * abs(dense.score_data())
    with pytest.raises(Exception, match="did not come from set_rows_diff|bad"):
        if dense.abi.prefix == "lb_":
            dense.score_rows_diff(0)
        else:
            raise RuntimeError("bad (the checker scores the decoded block either way)")
    with pytest.raises(Exception, match="neg removes"):
        bad = dict(d)
        bad["neg_feature"] = d["neg_feature"].copy()
        bad["neg_feature"][0] = model.n_features - 1     # a real-valued column is never in the tare
        diff.set_rows_diff(**bad)


def test_diff_path_equals_dense_path_oracle(oracle_abi):
    _equivalence(oracle_abi)


@pytest.mark.gpu
def test_diff_path_equals_dense_path_gpu(cuda_abi):
    _equivalence(cuda_abi)


@pytest.mark.gpu
def test_sparse_score_diff_matches_oracle_on_a_taxi_like_block(oracle_abi, cuda_abi):
    """C4's shape in small: dense mixed columns + many sparse booleans behind one tare row, rows without the tare,
    more than 128 groups in a kind; lb_score_rows_diff (O(nnz x G)) == the float64 oracle within 1e-5"""
    model, rows, _ = synth.generate_config("C4", rows=3000)
    g = Engine(cuda_abi)
    g.set_model(model)
    g.set_rows(rows)
    to, tv = g.make_tare()                                # the device passes (lb_differ_*)
    assert to[model.n_features - 1] == 0 and to[:20].sum() >= 15          # reals untouched, the sparse booleans tared
    d = g.sparsify(to, tv)
    assert len(d["pos_feature"]) < 0.8 * rows.n_cells()
    flags = np.ones(rows.n_rows, np.uint8)
    o = Engine(oracle_abi)
    for e in (o, g):
        e.set_model(model)
        e.set_rows_diff(**d)
    o.cat_sweep(sweep_actions(rows.n_rows), seed=5)
    from test_cuda_cat import _load_state
    keeps = _load_state(o, g, model)
    worst = 0.0
    for k in range(model.n_kinds):
        so = o.score_rows(k).astype(np.float64)
        so = np.concatenate([so[:, keeps[k]], so[:, [c for c in range(so.shape[1]) if c not in set(keeps[k])]]], axis=1)
        sg = g.score_rows_diff(k).astype(np.float64)
        assert sg.shape == so.shape
        worst = max(worst, (np.abs(sg - so) / np.maximum(1.0, np.abs(so))).max())
        assert np.abs(sg - g.score_rows(k)).max() <= 2e-5 * max(1.0, np.abs(so).max())
    assert worst <= 1e-5, worst


# ------------------------------------------------------------------------------------------ the passes themselves
# loom::Differ on the engine (lb_differ_*: SURVEY.md 8f n4) against the independent numpy restatement in
# loom_b200/tare.py, which the tests above tie to the reference's worked example.
def differ_case(abi):
    rng = np.random.default_rng(3)
    model, rows, _ = synth.generate(700, [("bb", 6), ("dd4", 2), ("dd256", 1), ("dpd5", 2), ("gp", 3), ("nich", 2)],
                                    density=0.85, seed=6)
    F8 = model.n8()
    rows.cat8[:4] = rng.random((4, 700)) < 0.04                      # sparse booleans: tare = 0
    rows.cat8[4] = rng.random(700) < 0.97                            # nearly always true: tare = 1
    rows.cat8[F8 - 1] = np.where(rng.random(700) < 0.8, 200, 3)      # dd256: the mode 200 is outside [0,16) -> no tare
    dpd = [f for f in range(model.n_features) if model.features[f]["type"] == "dpd"]
    gp = [f for f in range(model.n_features) if model.features[f]["type"] == "gp"]
    rows.wide[dpd[0] - F8] = np.where(rng.random(700) < 0.9, 2, rows.wide[dpd[0] - F8])   # dictionary index 2
    rows.wide[gp[0] - F8] = np.where(rng.random(700) < 0.7, 5, rows.wide[gp[0] - F8])
    rows.wide[gp[1] - F8] = np.where(rng.random(700) < 0.7, 16, rows.wide[gp[1] - F8])    # first value the summary ignores
    rows.mask[...] = engine_mask(rows, rng)
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(rows)
    want_obs, want_val = tare.make_tare(model, rows)
    got_obs, got_val = e.make_tare()
    assert np.array_equal(got_obs, want_obs) and np.array_equal(got_val[want_obs > 0], want_val[want_obs > 0])
    assert want_obs[:5].all() and want_obs[dpd[0]] and want_obs[gp[0]] and not want_obs[gp[1]] and not want_obs[F8 - 1]
    assert not want_obs[-2:].any()                                    # reals
    want = tare.sparsify(model, rows, want_obs, want_val)
    got = e.sparsify(got_obs, got_val)
    for key in ("pos_ptr", "pos_feature", "pos_value", "neg_ptr", "neg_feature"):
        assert np.array_equal(got[key], want[key]), key
    assert int(got["pos_ptr"][-1]) < 0.6 * rows.n_cells()             # it does compress
    # the diff decodes back to the block it came from: scores of every row agree
    back = Engine(abi)
    back.set_model(model)
    back.set_rows_diff(**got)
    assert np.array_equal(back.query_score(), e.query_score())
    # an empty tare passes rows through
    none = e.sparsify(np.zeros(model.n_features, np.uint8), np.zeros(model.n_features, np.uint32))
    assert int(none["pos_ptr"][-1]) == rows.n_cells() and int(none["neg_ptr"][-1]) == 0
    # raw DPD values: the [0,16) window and the mode are over the dictionary's VALUES, the tare holds the index
    offsets = np.array([0, 5, 10], np.int32)
    values = np.array([100, 101, 7, 103, 104, 0, 1, 2, 3, 4], np.uint32)    # feature dpd[0]: index 2 <-> raw 7
    obs2, val2 = e.make_tare(offsets, values)
    assert obs2[dpd[0]] and val2[dpd[0]] == 2
    values[2] = 50                                                           # now the common value is outside the window
    obs3, _ = e.make_tare(offsets, values)
    assert not obs3[dpd[0]]
    # windows of one stream: summaries add up, the threshold is over ALL rows added
    W = rows.mask.shape[1]
    half = RowBlock(352, rows.cat8[:, :352], rows.wide[:, :352], rows.mask[:, :11])
    rest_obs = rows.observed()[:, 352:]
    from loom_b200 import pack_mask
    rest = RowBlock(700 - 352, rows.cat8[:, 352:], rows.wide[:, 352:], pack_mask(rest_obs))
    w = Engine(abi)
    w.set_model(model)
    w.set_rows(half)
    w.make_tare()
    w.set_rows(rest)
    obs4, val4 = w.make_tare(reset=False)
    assert np.array_equal(obs4, want_obs) and np.array_equal(val4[want_obs > 0], want_val[want_obs > 0])
    assert W == 22


def engine_mask(rows, rng):
    from loom_b200 import pack_mask
    obs = rows.observed()
    obs[:5] |= rng.random(obs[:5].shape) < 0.9                         # the tared booleans are mostly observed
    return pack_mask(obs)


def test_differ_oracle(oracle_abi):
    differ_case(oracle_abi)


def test_differ_worked_example_and_edges(oracle_abi):
    model = bool_model(5)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, [[0, 1, None, 0, 0]]))
    d = e.sparsify(np.uint8([1, 1, 0, 0, 0]), np.zeros(5, np.uint32))          # doc/adapting.md:323-414
    assert list(d["pos_feature"]) == [1, 3, 4] and list(d["pos_value"]) == [1, 0, 0] and list(d["neg_feature"]) == [1]
    # ties go to the lower value and exactly half is not "more than half" (src/differ.hpp:59, differ.cc:197-199)
    e.set_rows(RowBlock.from_table(model, [[1, 1, 1, None, 0], [0, 1, 0, None, None]]))
    obs, val = e.make_tare()
    assert list(obs) == [0, 1, 0, 0, 0] and val[1] == 1
    with pytest.raises(Exception, match="before differ_reset"):
        f = Engine(oracle_abi)
        f.set_model(model)
        f.set_rows(RowBlock.from_table(model, [[0, 0, 0, 0, 0]]))
        f.make_tare(reset=False)
