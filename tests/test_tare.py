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
    to, tv = tare.make_tare(model, rows)
    assert to[:4].all()
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
    assert dense.score_data() == diff.score_data()
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
