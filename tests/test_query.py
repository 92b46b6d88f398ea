"""Query sampling (the per-latent half of QueryServer::call(Sample), src/query_server.cc:100-176):
the oracle's draws against its own scores -- the reference's sample / score self-consistency check
(loom/test/test_query_math.py:65-92) -- per feature type, jointly inside a kind, and across kinds.
The GPU side of the same checks is in tests/test_z_loom_infer.py."""
import numpy as np
import pytest
from scipy import stats

from loom_b200 import Engine, RowBlock, sweep_actions, synth
from loom_b200._abi import DPD_OTHER, UNASSIGNED

TYPES = [("bb", 4), ("dd4", 3), ("dd16", 2), ("dpd5", 2), ("gp", 3), ("nich", 3)]
N_DRAWS = 20000
MIN_P = 1e-3          # chi-square / KS p-values must exceed this (draws are deterministic: no flakiness)


def trained(abi, rows=400, seed=11, f2k=None):
    model, block, _ = synth.generate(rows, TYPES, density=0.7, seed=seed, f2k=f2k)
    e = Engine(abi)
    e.set_model(model)
    e.set_rows(block)
    e.cat_sweep(np.concatenate([sweep_actions(rows), sweep_actions(rows, add=True, remove=True)]), seed=3)
    return model, e


def frozen(abi, model, source, table):
    """a latent sample as the query server holds it: `source`'s groups, no assignments, `table` as rows"""
    q = Engine(abi)
    q.set_model(model)
    q.set_rows(RowBlock.from_table(model, table))
    for k in range(model.n_kinds):
        counts, ints, flts = source.get_groups(k)
        keep = counts > 0
        q.set_groups(k, int(keep.sum()), counts[keep], ints[:, keep], flts[:, keep])
    return q


def conditioning_row(model, observe):
    row = [None] * model.n_features
    for f in observe:
        t = model.features[f]["type"]
        row[f] = {"bb": 1, "dd": 1, "dpd": 2, "gp": 3, "nich": 0.25}[t]
    return row


def candidates(model, f):
    t = model.features[f]["type"]
    if t == "bb":
        return [0, 1]
    if t == "dd":
        return list(range(len(model.features[f]["alphas"])))
    if t == "dpd":
        return list(range(len(model.features[f]["betas"])))
    if t == "gp":
        return list(range(80))
    raise ValueError(t)


def predictive(abi, model, source, cond, cells_list):
    """p(cells | cond) for every dict {feature: value} of cells_list, from query_score (Score)"""
    table = [cond]
    for cells in cells_list:
        row = list(cond)
        for f, v in cells.items():
            row[f] = v
        table.append(row)
    s = frozen(abi, model, source, table).query_score().astype(np.float64)
    return np.exp(s[1:] - s[0])


def chi_square_p(counts, probs):
    """goodness of fit with small cells merged into one"""
    counts, probs = np.asarray(counts, float), np.asarray(probs, float)
    n = counts.sum()
    small = probs * n < 5
    if small.sum() > 1:
        counts = np.append(counts[~small], counts[small].sum())
        probs = np.append(probs[~small], probs[small].sum())
    probs = probs / probs.sum()
    return stats.chisquare(counts, probs * n).pvalue


@pytest.fixture(scope="module")
def setup(oracle_abi):
    model, e = trained(oracle_abi)
    observe = [f for f in range(model.n_features) if f % 3 == 0]
    cond = conditioning_row(model, observe)
    to_sample = [f for f in range(model.n_features) if f % 3 != 0]
    q = frozen(oracle_abi, model, e, [cond])
    values, groups = q.query_sample(0, to_sample, N_DRAWS, seed=7, groups=True)
    return model, e, cond, to_sample, values, groups


def test_groups_are_drawn_for_exactly_the_kinds_sampled(setup):
    model, e, cond, to_sample, values, groups = setup
    wanted = sorted({int(model.f2k[f]) for f in to_sample})
    for k in range(model.n_kinds):
        if k in wanted:
            assert (groups[:, k] < e.kind_info(k).n_groups).all()
        else:
            assert (groups[:, k] == UNASSIGNED).all()


def test_group_frequencies_follow_the_conditional_posterior(oracle_abi, setup):
    model, e, cond, to_sample, values, groups = setup
    q = frozen(oracle_abi, model, e, [cond])
    for k in sorted({int(model.f2k[f]) for f in to_sample}):
        s = q.score_rows(k)[0].astype(np.float64)
        p = np.exp(s - s.max())
        p /= p.sum()
        assert chi_square_p(np.bincount(groups[:, k], minlength=len(p)), p) > MIN_P


@pytest.mark.parametrize("ftype", ["bb", "dd", "dpd", "gp"])
def test_discrete_marginals_match_scores(oracle_abi, setup, ftype):
    model, e, cond, to_sample, values, groups = setup
    checked = 0
    for j, f in enumerate(to_sample):
        if model.features[f]["type"] != ftype:
            continue
        cand = candidates(model, f)
        p = predictive(oracle_abi, model, e, cond, [{f: v} for v in cand])
        col = values[:, j]
        counts = np.array([(col == v).sum() for v in cand] + [0.0])
        counts[-1] = len(col) - counts[:-1].sum()          # DPD: unseen value; GP: the tail
        rest = max(1.0 - p.sum(), 0.0)
        if ftype in ("bb", "dd"):
            assert abs(p.sum() - 1) < 1e-5 and counts[-1] == 0
            assert chi_square_p(counts[:-1], p) > MIN_P
        else:
            if ftype == "dpd":
                assert ((col == DPD_OTHER) | (col < len(cand))).all()
                assert rest > 1e-3                           # the unseen mass alpha * beta0 / (alpha + N) is visible here
            assert chi_square_p(counts, np.append(p, rest)) > MIN_P
        checked += 1
    assert checked >= 1


def test_nich_marginals_match_scores(oracle_abi, setup):
    model, e, cond, to_sample, values, groups = setup
    checked = 0
    for j, f in enumerate(to_sample):
        if model.features[f]["type"] != "nich":
            continue
        x = values[:, j].copy().view(np.float32).astype(np.float64)
        assert np.isfinite(x).all()
        lo, hi = np.quantile(x, [0.001, 0.999])
        grid = np.linspace(lo - 5 * (hi - lo), hi + 5 * (hi - lo), 6001)
        dens = predictive(oracle_abi, model, e, cond, [{f: float(np.float32(v))} for v in grid])
        cdf = np.concatenate([[0], np.cumsum(0.5 * (dens[1:] + dens[:-1]) * np.diff(grid))])
        assert 0.97 < cdf[-1] <= 1.001                      # heavy tails (nu' small) leave a little mass outside
        cdf /= cdf[-1]
        inside = (x > grid[0]) & (x < grid[-1])
        assert inside.mean() > 0.97
        d = stats.kstest(x[inside], lambda t: np.interp(t, grid, cdf))
        assert d.pvalue > MIN_P, (f, d)
        checked += 1
    assert checked >= 1


def test_features_of_one_kind_share_the_group(oracle_abi, setup):
    """joint law of two features of the same kind = mixture over ONE group draw, not a product of marginals"""
    model, e, cond, to_sample, values, groups = setup
    pairs = [(a, b) for a in to_sample for b in to_sample
             if a < b and model.f2k[a] == model.f2k[b]
             and model.features[a]["type"] in ("bb", "dd") and model.features[b]["type"] in ("bb", "dd")]
    assert pairs
    a, b = pairs[0]
    ca, cb = candidates(model, a), candidates(model, b)
    cells = [{a: x, b: y} for x in ca for y in cb]
    p = predictive(oracle_abi, model, e, cond, cells)
    assert abs(p.sum() - 1) < 1e-5
    ja, jb = to_sample.index(a), to_sample.index(b)
    counts = [((values[:, ja] == c[a]) & (values[:, jb] == c[b])).sum() for c in cells]
    assert chi_square_p(counts, p) > MIN_P


def test_draws_are_a_function_of_seed_and_offset(oracle_abi, setup):
    model, e, cond, to_sample, values, groups = setup
    q = frozen(oracle_abi, model, e, [cond])
    again = q.query_sample(0, to_sample, 100, seed=7, draw_offset=500)
    assert np.array_equal(again, values[500:600])
    other = q.query_sample(0, to_sample, 100, seed=8)
    assert not np.array_equal(other, values[:100])
    assert q.query_sample(0, [], 5, seed=1).shape == (5, 0)
    with pytest.raises(Exception, match="ascending"):
        q.query_sample(0, [3, 2], 1, seed=1)
    with pytest.raises(Exception, match="bad feature"):
        q.query_sample(0, [model.n_features], 1, seed=1)
    with pytest.raises(Exception, match="bad row"):
        q.query_sample(1, [0], 1, seed=1)


def test_large_poisson_rates_use_the_rejection_sampler(oracle_abi):
    """a GP group with a large sum: the PTRS branch (rate >= 30) against the negative-binomial law"""
    model, block, _ = synth.generate(50, [("gp", 1)], density=1.0, seed=2, single_kind=True)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, [[None]]))
    # one group: count 40, sum 4000 -> posterior rate ~ Gamma(4002, 1/40.5), mean ~ 98.8
    e.set_groups(0, 1, np.array([40], np.uint32), np.array([[40], [4000]], np.uint32), np.array([[0.0]]))
    g0 = e.query_sample(0, [0], N_DRAWS, seed=5, groups=True)
    x = g0[0][:, 0].astype(np.float64)
    picked = g0[1][:, 0]
    s = e.score_rows(0)[0].astype(np.float64)
    p_group = np.exp(s - s.max())
    p_group /= p_group.sum()
    assert chi_square_p(np.bincount(picked, minlength=2), p_group) > MIN_P
    xs = x[picked == 0]
    r, prob = 2.0 + 4000, 40.5 / 41.5
    edges = np.arange(60, 140, 4)
    counts = np.histogram(xs, np.concatenate([[-1], edges, [1e9]]))[0]
    cdf = stats.nbinom.cdf(np.concatenate([edges - 1, [1e9]]), r, prob)      # bins are [a, b)
    probs = np.diff(np.concatenate([[0], cdf]))
    assert chi_square_p(counts, probs) > MIN_P


def test_small_gamma_shapes_use_the_boost(oracle_abi):
    """a GP feature with alpha < 1 and no data: the prior predictive NB(alpha, inv_beta / (inv_beta + 1))"""
    from loom_b200 import Model
    model = Model([dict(type="gp", alpha=0.3, inv_beta=0.5)], [0])
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(RowBlock.from_table(model, [[None]]))
    x = e.query_sample(0, [0], N_DRAWS, seed=9)[:, 0].astype(np.int64)
    top = 12
    counts = np.append(np.bincount(np.minimum(x, top), minlength=top + 1)[:top], (x >= top).sum())
    pmf = stats.nbinom.pmf(np.arange(top), 0.3, 0.5 / 1.5)
    assert chi_square_p(counts, np.append(pmf, 1 - pmf.sum())) > MIN_P
