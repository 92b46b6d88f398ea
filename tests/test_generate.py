"""generate_rows (Loom::generate -> generate_rows, src/generate.hpp:36-93) on the oracle: the engine's books balance
(the statistics it ends with are those of the rows it wrote under the assignments it made), the law of a first row is
the prior predictive, the seating follows the Pitman-Yor process, and a generated dataset is recovered by inference
no worse than it scores under its own generating state.  The device against the oracle: tests/test_zz_query_gpu.py."""
import numpy as np
import pytest
from scipy import stats

from loom_b200 import Engine, Model, synth
from test_query import MIN_P, chi_square_p

TYPES = [("bb", 3), ("dd4", 2), ("dd16", 1), ("dpd5", 2), ("gp", 2), ("nich", 2)]


def generated(abi, n_rows=400, density=0.7, seed=3, model=None):
    if model is None:
        model, _, _ = synth.generate(10, TYPES, density=1.0, seed=1)
    e = Engine(abi)
    e.set_model(model)
    e.generate_rows(n_rows, density, seed)
    return model, e


def books_balance(abi, model, e):
    """recompute every statistic from the generated rows and assignments"""
    rows = e.get_rows()
    check = Engine(abi)
    check.set_model(model)
    check.set_rows(rows)
    for k in range(model.n_kinds):
        a = e.get_assignments(k)
        assert (a != 0xFFFFFFFF).all()
        counts, ints, flts = e.get_groups(k)
        n = int(np.count_nonzero(counts))
        assert (counts[:n] > 0).all() and counts[:n].sum() == rows.n_rows          # seated groups first, then the empty one
        assert np.array_equal(np.bincount(a, minlength=len(counts)), counts)
        check.set_groups(k, n)
        check.set_assignments(k, a)
    check.accumulate_rows()
    for k in range(model.n_kinds):
        (c0, i0, f0), (c1, i1, f1) = e.get_groups(k), check.get_groups(k)
        assert np.array_equal(c0, c1) and np.array_equal(i0, i1)
        assert np.allclose(f0, f1, rtol=1e-9, atol=1e-9)
    assert e.score_data() == pytest.approx(check.score_data(), rel=1e-9)
    return rows


def test_books_balance_and_density(oracle_abi):
    model, e = generated(oracle_abi)
    rows = books_balance(oracle_abi, model, e)
    obs = rows.observed()
    # observed ~ Bernoulli(density) per cell, except DPD draws outside the dictionary (left unobserved)
    plain = [f for f in range(model.n_features) if model.features[f]["type"] != "dpd"]
    n = obs[plain].size
    assert abs(obs[plain].mean() - 0.7) < 4 * np.sqrt(0.21 / n)
    _, e0 = generated(oracle_abi, density=0.0)
    assert e0.get_rows().n_cells() == 0
    _, e1 = generated(oracle_abi, density=1.0)
    assert e1.get_rows().observed()[plain].all()
    model2, again = generated(oracle_abi)
    assert np.array_equal(again.get_rows().cat8, rows.cat8) and np.array_equal(again.get_rows().wide, rows.wide)
    _, other = generated(oracle_abi, seed=4)
    assert not np.array_equal(other.get_rows().cat8, rows.cat8)
    with pytest.raises(Exception, match="density"):
        e.generate_rows(10, 1.5, 0)


def test_first_row_follows_the_prior_predictive(oracle_abi):
    """many one-row datasets: each cell of the first row is a draw from the feature's prior predictive"""
    model = Model([dict(type="bb", alpha=0.7, beta=1.9), dict(type="dd", alphas=[0.5, 1.0, 2.5]),
                   dict(type="gp", alpha=2.5, inv_beta=0.8), dict(type="nich", mu=1.0, kappa=2.0, sigmasq=1.5, nu=5.0)], [0, 0, 1, 1])
    e = Engine(oracle_abi)
    e.set_model(model)
    n = 3000
    cells = np.zeros((4, n))
    for s in range(n):
        e.generate_rows(1, 1.0, 1000 + s)
        r = e.get_rows()
        cells[0, s], cells[1, s] = r.cat8[0, 0], r.cat8[1, 0]
        cells[2, s], cells[3, s] = r.wide[0, 0], r.wide[1, :1].view(np.float32)[0]
    ones = cells[0].sum()
    assert chi_square_p([n - ones, ones], [1.9 / 2.6, 0.7 / 2.6]) > MIN_P
    assert chi_square_p(np.bincount(cells[1].astype(int), minlength=3), np.array([0.5, 1.0, 2.5]) / 4.0) > MIN_P
    top = 14
    counts = np.append(np.bincount(np.minimum(cells[2].astype(int), top), minlength=top + 1)[:top], (cells[2] >= top).sum())
    pmf = stats.nbinom.pmf(np.arange(top), 2.5, 0.8 / 1.8)
    assert chi_square_p(counts, np.append(pmf, 1 - pmf.sum())) > MIN_P
    scale = np.sqrt(1.5 * (2.0 + 1.0) / 2.0)
    assert stats.kstest(cells[3], lambda x: stats.t.cdf((x - 1.0) / scale, 5.0)).pvalue > MIN_P


def test_seating_follows_the_pitman_yor_process(oracle_abi):
    """number of groups among n rows under PY(alpha, d): the mean over many seeds against the exact expectation"""
    alpha, d, n = 2.0, 0.1, 60
    model = Model([dict(type="bb", alpha=1.0, beta=1.0)], [0], kind_py=[(alpha, d)])
    e = Engine(oracle_abi)
    e.set_model(model)
    tables = []
    for s in range(600):
        e.generate_rows(n, 0.5, 50 + s)
        tables.append(int(np.count_nonzero(e.get_groups(0)[0])))
    # E[K_n] = sum_i P(row i opens a group), with E[alpha + d K_i] propagated exactly
    ek, expect = 0.0, 0.0
    for i in range(n):
        p_new = (alpha + d * ek) / (alpha + i)
        expect += p_new
        ek += p_new
    assert abs(np.mean(tables) - expect) < 4 * np.std(tables) / np.sqrt(len(tables))


def test_inference_recovers_a_generated_dataset(oracle_abi):
    """a Gibbs chain started from scratch on generated rows reaches a data score in the range of the generating state's"""
    from loom_b200 import sweep_actions
    model, e = generated(oracle_abi, n_rows=300, density=0.9, seed=8)
    truth = e.score_data()
    rows = e.get_rows()
    chain = Engine(oracle_abi)
    chain.set_model(model)
    chain.set_rows(rows)
    acts = [sweep_actions(300)] + [sweep_actions(300, add=True, remove=True)] * 12
    chain.cat_sweep(np.concatenate(acts), seed=5)
    assert chain.score_data() > truth - 0.1 * abs(truth)
