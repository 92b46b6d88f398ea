"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py).

CPU: the oracle against scipy's known answers for every closed form, the reference's structural
fixtures, and the oracle's own recorded results on a small mixed model (regression).
GPU: the CUDA path against the same recorded results, through the C ABI.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from loom_b200 import Engine, Model, RowBlock
from loom_b200._abi import FeatureSpec, TYPE_NAMES
import pe_util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCORE_TOL = 1e-5   # north star: scores within 1e-5 relative (|a-b| <= tol * max(1, |b|))


def rel_err(a, b):
    return np.abs(np.asarray(a, np.float64) - b) / np.maximum(1.0, np.abs(b))


@pytest.fixture(scope="module")
def small():
    z = np.load(os.path.join(GOLDEN, "engine_small.npz"))
    model = Model(json.loads(str(z["features"])), z["f2k"], kind_py=z["kind_py"])
    rows = RowBlock(int(z["n_rows"]), z["cat8"], z["wide"], z["mask"])
    return z, model, rows


def _bind(lib):
    lib.lo_feature_score_value.restype = C.c_double
    lib.lo_feature_score_value.argtypes = [C.POINTER(FeatureSpec), C.POINTER(C.c_float),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_uint32]
    lib.lo_feature_score_data.restype = C.c_double
    lib.lo_feature_score_data.argtypes = lib.lo_feature_score_value.argtypes[:4]
    lib.lo_py_score_counts.restype = C.c_double
    lib.lo_py_score_counts.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_uint32), C.c_int]
    return lib


def test_oracle_closed_forms_match_scipy_fixture(oracle_abi):
    lib = _bind(oracle_abi.lib)
    with open(os.path.join(GOLDEN, "closed_forms.json")) as f:
        kat = json.load(f)
    assert len(kat["features"]) >= 50
    for c in kat["features"]:
        s = FeatureSpec()
        s.type, s.dim, s.aux_off = TYPE_NAMES[c["type"]], c.get("dim", 0), 0
        for i, v in enumerate(c.get("hp", [])):
            s.hp[i] = v
        aux = np.asarray(c.get("aux") or [0], np.float32)
        ints = np.asarray(c["ints"], np.uint32)
        flts = np.asarray(c["flts"] or [0], np.float64)
        args = (C.byref(s), aux.ctypes.data_as(C.POINTER(C.c_float)), ints.ctypes.data_as(C.POINTER(C.c_uint32)),
                flts.ctypes.data_as(C.POINTER(C.c_double)))
        for v, want in zip(c["values"], c["predictive"]):
            got = lib.lo_feature_score_value(*args, v)
            assert got == pytest.approx(want, rel=1e-9, abs=1e-9), (c["type"], v)
        got = lib.lo_feature_score_data(*args)
        assert got == pytest.approx(c["marginal"], rel=1e-9, abs=1e-8), c["type"]
    for c in kat["pitman_yor"]:
        counts = np.asarray(c["counts"], np.uint32)
        got = lib.lo_py_score_counts(c["alpha"], c["d"], counts.ctypes.data_as(C.POINTER(C.c_uint32)), len(counts))
        assert got == pytest.approx(c["score_counts"], rel=1e-10, abs=1e-10)


def test_structural_fixtures_of_the_reference():
    with open(os.path.join(GOLDEN, "structural.json")) as f:
        st = json.load(f)
    # LATENT_SIZES[rows][cols] counts cross-cat latent states: Bell numbers in row/column 1,
    # count_crosscats in general (loom/test/test_posterior_enum.py:65-79,791-794)
    sizes, bell = st["LATENT_SIZES"]["value"], st["BELL_NUMBERS"]["value"]
    assert sizes[1] == bell[:len(sizes[1])]
    for r in range(1, 5):
        for c in range(1, min(5, len(sizes[r]))):
            assert pe_util.count_crosscats(r, c) == sizes[r][c]
    # the worked tare example (doc/adapting.md:323-414) through sparsify
    from loom_b200 import tare
    ex = st["tare_example"]
    model = Model([dict(type="bb", alpha=0.5, beta=0.5)] * 5, [0] * 5)
    rows = RowBlock.from_table(model, [[None if v is None else int(v) for v in ex["row"]]])
    d = tare.sparsify(model, rows, np.uint8(ex["tare"]["observed"]), np.zeros(5, np.uint32))
    pos_obs, neg_obs = ex["diff"]["pos"]["observed"], ex["diff"]["neg"]["observed"]
    assert list(d["pos_feature"]) == [i for i, o in enumerate(pos_obs) if o]
    assert [bool(v) for v in d["pos_value"]] == ex["diff"]["pos"]["booleans"]
    assert list(d["neg_feature"]) == [i for i, o in enumerate(neg_obs) if o]


def _check_fixed(abi, small, int_exact, score_tol):
    z, model, rows = small
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(rows)
    for k in range(model.n_kinds):
        a = z["fixed_assign_%d" % k]
        e.set_groups(k, int(a.max()) + 1)
        e.set_assignments(k, a)
    e.accumulate_rows()
    for k in range(model.n_kinds):
        c, i, f = e.get_groups(k)
        assert np.array_equal(c, z["fixed_counts_%d" % k])
        assert np.array_equal(i, z["fixed_ints_%d" % k])          # integer statistics: bit-exact
        assert np.allclose(f, z["fixed_flts_%d" % k], rtol=1e-9 if int_exact else 1e-9, atol=1e-9)
        s = e.score_rows(k)[:64]
        assert rel_err(s, z["fixed_scores_%d" % k]).max() <= score_tol
    assert rel_err(e.score_data(), float(z["fixed_score_data"])) <= score_tol
    assert rel_err(e.kind_proposer_score(), z["fixed_L"]).max() <= score_tol
    return e


def test_oracle_reproduces_recorded_results(oracle_abi, small):
    z, model, rows = small
    _check_fixed(oracle_abi, small, True, 1e-12)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    picks, _ = e.cat_sweep(z["sweep_actions"], seed=int(z["sweep_seed"]), debug=True, debug_stride=8)
    assert np.array_equal(picks, z["sweep_picks"])
    for k in range(model.n_kinds):
        assert np.array_equal(e.get_assignments(k), z["sweep_assign_%d" % k])
        c, i, _f = e.get_groups(k)
        assert np.array_equal(c, z["sweep_counts_%d" % k]) and np.array_equal(i, z["sweep_ints_%d" % k])
    assert e.score_data() == pytest.approx(float(z["sweep_score_data"]), rel=1e-12)


@pytest.mark.gpu
def test_cuda_reproduces_recorded_results(cuda_abi, small):
    z, model, rows = small
    _check_fixed(cuda_abi, small, False, SCORE_TOL)
    # the recorded sweep's final state, installed as given: same statistics, same data score
    e = Engine(cuda_abi)
    e.set_model(model)
    e.set_rows(rows)
    used = []
    for k in range(model.n_kinds):
        uniq, inv = np.unique(z["sweep_assign_%d" % k], return_inverse=True)   # drop the empty group(s)
        used.append(uniq)
        e.set_groups(k, len(uniq))
        e.set_assignments(k, inv.astype(np.uint32))
    e.accumulate_rows()
    for k in range(model.n_kinds):
        c, i, _f = e.get_groups(k)
        want_c, want_i = z["sweep_counts_%d" % k], z["sweep_ints_%d" % k]
        n = len(used[k])                     # get_groups also returns the trailing empty group
        assert np.array_equal(c[:n], want_c[used[k]]) and want_c.sum() == c.sum()
        assert np.array_equal(i[:, :n], want_i[:, used[k]]) and not i[:, n:].any()
    assert rel_err(e.score_data(), float(z["sweep_score_data"])) <= SCORE_TOL


def test_hyperprior_grids_match_the_reference_gridding():
    """loom_b200/hyperprior.py against the outputs of the reference's own loom/gridding.py
    (tests/golden/reference_grids.json) and the grid sizes of loom/hyperprior.py:33-71."""
    from loom_b200 import hyperprior as hp
    with open(os.path.join(GOLDEN, "reference_grids.json")) as f:
        ref = json.load(f)
    assert np.allclose(hp.uniform(0, 1, 7), ref["uniform_0_1_7"], rtol=1e-12)
    assert np.allclose(hp.center_heavy(0, 1, 20), ref["center_heavy_0_1_20"], rtol=1e-12)
    assert np.allclose(hp.left_heavy(-1, 2, 30), ref["left_heavy_m1_2_30"], rtol=1e-12)
    assert np.allclose(hp.right_heavy(-1, 1, 20), ref["right_heavy_m1_1_20"], rtol=1e-12)
    mine = hp.pitman_yor()
    assert len(mine) == len(ref["pitman_yor"])
    for a, b in zip(mine, ref["pitman_yor"]):
        assert a["alpha"] == pytest.approx(b["alpha"], rel=1e-12) and a["d"] == pytest.approx(b["d"], rel=1e-12, abs=1e-15)
    d = hp.defaults()
    assert len(d["dd"]["alpha"]) == 12 and len(d["gp"]["alpha"]) == 100 and len(d["nich"]["mu"]) == 201
    assert len(d["nich"]["kappa"]) == 30 and len(d["dpd"]["gamma"]) == 30 and len(d["dpd"]["alpha"]) == 20


# ------------------------------------------------------------------------------------------ query / differ fixture
def query_differ_case(abi, small, exact):
    """tests/golden/query_differ_small.npz: Score, a seeded batch of draws, the tare row and a diff of the block.
    exact = the oracle itself (regression, bit for bit); otherwise the CUDA path (scores to 1e-5, the diff bit for
    bit, draws wherever the group pick agrees)."""
    z, model, rows = small
    q = np.load(os.path.join(GOLDEN, "query_differ_small.npz"))
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(rows)
    for k in range(model.n_kinds):
        a = z["fixed_assign_%d" % k]
        e.set_groups(k, int(a.max()) + 1)
        e.set_assignments(k, a)
    e.accumulate_rows()
    assert rel_err(e.query_score(0, 64), q["query_scores"]).max() <= (1e-7 if exact else SCORE_TOL)
    values, groups = e.query_sample(0, q["to_sample"], 256, seed=int(q["sample_seed"]), groups=True)
    if exact:
        assert np.array_equal(values, q["sample_values"]) and np.array_equal(groups, q["sample_groups"])
    else:
        assert (groups == q["sample_groups"]).mean() > 0.995
        for j, f in enumerate(q["to_sample"]):
            same = groups[:, model.f2k[f]] == q["sample_groups"][:, model.f2k[f]]
            a, b = values[same, j], q["sample_values"][same, j]
            if model.features[int(f)]["type"] == "nich":
                a, b = a.copy().view(np.float32).astype(np.float64), b.copy().view(np.float32).astype(np.float64)
                assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
            else:
                assert (a == b).mean() > 0.99
    obs, val = e.make_tare()
    assert np.array_equal(obs, q["tare_observed"]) and np.array_equal(val, q["tare_values"])
    d = e.sparsify(q["hand_tare_observed"], q["hand_tare_values"])
    for key in ("pos_ptr", "pos_feature", "pos_value", "neg_ptr", "neg_feature"):
        assert np.array_equal(d[key], q[key]), key


def test_oracle_reproduces_query_and_differ_fixture(oracle_abi, small):
    query_differ_case(oracle_abi, small, True)
