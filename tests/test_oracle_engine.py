"""Oracle self-checks: the sequential cat kernel keeps its invariants, bulk and
sequential sufficient statistics agree, and the sampled assignments pass loom's
posterior-enumeration chi-square test (the only distributional anchor the
reference offers: loom/test/test_posterior_enum.py)."""
import numpy as np
import pytest

import pe_util
from loom_b200 import Engine, sweep_actions, UNASSIGNED
from loom_b200 import synth


def small_mixed(rows=300, seed=5):
    return synth.generate(rows, [("bb", 4), ("dd4", 3), ("dd16", 2), ("dpd5", 2), ("gp", 3),
                                 ("nich", 3)], density=0.6, seed=seed)


def test_add_remove_roundtrip(oracle_abi):
    model, rows, _ = small_mixed()
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    e.cat_sweep(sweep_actions(rows.n_rows), seed=1)
    for k in range(model.n_kinds):
        info = e.kind_info(k)
        assert info.sample_size == rows.n_rows
        counts, ints, flts = e.get_groups(k)
        assert counts.sum() == rows.n_rows
        assert (counts == 0).sum() == model.empty_group_count  # A5 invariant
        a = e.get_assignments(k)
        assert np.array_equal(np.bincount(a, minlength=len(counts)), counts)
    s1 = e.score_data()
    assert np.isfinite(s1)
    # remove everything: back to the empty state
    e.cat_sweep(sweep_actions(rows.n_rows, add=False), seed=2)
    for k in range(model.n_kinds):
        counts, ints, flts = e.get_groups(k)
        assert len(counts) == model.empty_group_count and counts.sum() == 0
        assert not ints.any()
        assert np.allclose(flts, 0, atol=1e-9)
        assert (e.get_assignments(k) == UNASSIGNED).all()


def test_multi_chain_sweep_and_shared_rows(oracle_abi):
    """cat_sweep_multi == one cat_sweep per chain with its own seed; share_rows aliases the owner's
    block (loom's samples read one rows file, loom/tasks.py:173-194)."""
    from loom_b200 import cat_sweep_multi
    model, rows, _ = synth.generate(300, [("bb", 3), ("dd4", 2), ("gp", 2), ("nich", 2)], density=0.7, seed=3)
    acts = np.concatenate([sweep_actions(300), sweep_actions(300, add=True, remove=True)])
    seeds = [5, 6, 7]
    singles = []
    for seed in seeds:
        e = Engine(oracle_abi)
        e.set_model(model)
        e.set_rows(rows)
        e.cat_sweep(acts, seed=seed)
        singles.append(e)
    owner = Engine(oracle_abi)
    owner.set_model(model)
    owner.set_rows(rows)
    chains = [owner]
    for _ in seeds[1:]:
        e = Engine(oracle_abi)
        e.set_model(model)
        e.share_rows(owner)
        chains.append(e)
    cat_sweep_multi(chains, acts, seeds)
    for a, b in zip(singles, chains):
        for k in range(model.n_kinds):
            assert np.array_equal(a.get_assignments(k), b.get_assignments(k))
            assert all(np.array_equal(x, y) for x, y in zip(a.get_groups(k), b.get_groups(k)))
    other = Engine(oracle_abi)
    other.set_model(synth.generate(10, [("bb", 2)], seed=1)[0])
    with pytest.raises(Exception, match="schemas differ"):
        other.share_rows(owner)
    with pytest.raises(Exception, match="borrows"):
        other2 = Engine(oracle_abi)
        other2.set_model(model)
        other2.share_rows(chains[1])


def test_query_score_is_a_normalised_predictive(oracle_abi):
    """QueryServer Score (src/query_server.cc:189-231): per sample, sum over kinds of
    log_sum_exp(score_value); over samples, log-mean-exp.  Summed over every possible row it is a
    probability distribution."""
    from loom_b200 import query_score, RowBlock
    model, rows, truth = synth.generate(200, [("bb", 2), ("dd4", 1)], density=1.0, seed=6, f2k=[0, 1, 0])
    chains = []
    for seed in (1, 2):
        e = Engine(oracle_abi)
        e.set_model(model)
        e.set_rows(rows)
        e.cat_sweep(sweep_actions(200), seed=seed)
        chains.append(e)
    # per sample: equals the sum over kinds of log-sum-exp of score_rows
    e = chains[0]
    want = sum(np.log(np.exp(e.score_rows(k).astype(np.float64)).sum(axis=1)) for k in range(model.n_kinds))
    assert np.allclose(e.query_score(), want, rtol=1e-5, atol=1e-5)
    # all 2 x 2 x 4 complete rows + normalisation, with the trained groups kept and the query rows loaded
    table = [[a, b, c] for a in (0, 1) for b in (0, 1) for c in range(4)]
    q = RowBlock.from_table(model, table)
    states = [[e.get_groups(k) for k in range(model.n_kinds)] for e in chains]
    for e, st in zip(chains, states):
        e.set_rows(q)
        for k, (c, i, f) in enumerate(st):
            e.set_groups(k, len(c), c, i, f)
    total = np.exp(query_score(chains)).sum()
    assert total == pytest.approx(1.0, rel=1e-5)
    # partially observed rows marginalise: p(a) = sum_b,c p(a, b, c)
    part = RowBlock.from_table(model, [[0, None, None], [1, None, None]])
    for e, st in zip(chains, states):
        e.set_rows(part)
        for k, (c, i, f) in enumerate(st):
            e.set_groups(k, len(c), c, i, f)
    assert np.exp(query_score(chains)).sum() == pytest.approx(1.0, rel=1e-5)


def test_duplicate_row_is_an_error(oracle_abi):
    model, rows, _ = small_mixed(50)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    e.cat_sweep(sweep_actions(10), seed=1)
    with pytest.raises(Exception, match="duplicate row"):  # cat_kernel.hpp:184
        e.cat_sweep(sweep_actions(1), seed=1)


def test_bulk_accumulate_matches_sequential(oracle_abi):
    model, rows, _ = small_mixed()
    a = Engine(oracle_abi)
    a.set_model(model)
    a.set_rows(rows)
    a.cat_sweep(sweep_actions(rows.n_rows), seed=3)
    a.cat_sweep(sweep_actions(rows.n_rows, add=True, remove=True), seed=4)
    b = Engine(oracle_abi)
    b.set_model(model)
    b.set_rows(rows)
    uniqs = []
    for k in range(model.n_kinds):
        uniq, inv = np.unique(a.get_assignments(k), return_inverse=True)
        uniqs.append(uniq)
        b.set_groups(k, len(uniq))
        b.set_assignments(k, inv.astype(np.uint32))
    b.accumulate_rows()
    for k in range(model.n_kinds):
        ca, ia, fa = a.get_groups(k)
        cb, ib, fb = b.get_groups(k)
        u, n = uniqs[k], len(uniqs[k])
        assert np.array_equal(ca[u], cb[:n]) and np.array_equal(ia[:, u], ib[:, :n])  # bit-exact
        assert np.allclose(fa[:, u], fb[:, :n], rtol=1e-9, atol=1e-9)
    assert a.score_data() == pytest.approx(b.score_data(), rel=1e-10)
    sa, sb = a.score_rows(0), b.score_rows(0)
    assert np.allclose(sa[:, uniqs[0]], sb[:, :len(uniqs[0])], rtol=1e-6, atol=1e-6)


def test_score_rows_is_normalizable_predictive(oracle_abi):
    # For a single fully-observed BB feature the two possible rows must have
    # total predictive mass 1 once the CRP prior is marginalised.
    model = pe_util.make_model(1, "bb")
    table = [[1], [0], [1], [1], [0], [1]]
    rows = pe_util.RowBlock.from_table(model, table)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    e.cat_sweep(sweep_actions(4), seed=9)
    s = e.score_rows(0, 4, 6).astype(np.float64)  # rows 4 (x=0) and 5 (x=1), unassigned
    total = np.exp(s).sum()
    assert total == pytest.approx(1.0, rel=1e-6)


@pytest.mark.parametrize("feature_type", pe_util.FEATURE_TYPES)
@pytest.mark.parametrize("density", pe_util.DENSITIES)
def test_posterior_enum_cats(oracle_abi, feature_type, density):
    """infer_cats (test_posterior_enum.py:139-160) at nose size 100, on the
    oracle itself -- validates the restated math + control flow."""
    rng = np.random.default_rng(123)
    for (R, C) in pe_util.cat_dimensions(100):
        model = pe_util.make_model(C, feature_type)
        rows = pe_util.RowBlock.from_table(
            model, pe_util.generate_rows(R, C, feature_type, density, rng))
        e = Engine(oracle_abi)
        samples = pe_util.run_posterior_enum(e, model, rows, 10 * pe_util.LATENT_SIZES[R][1],
                                             seed=R * 100 + C)
        gof, info = pe_util.goodness_of_fit(samples, R, C, infer_kinds=False)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)
