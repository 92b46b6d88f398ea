"""Posterior-enumeration (Geweke-style) chi-square test, restated from
loom/test/test_posterior_enum.py for Python 3 and driven through the C ABI
instead of the loom_posterior_enum binary.

Protocol (src/loom.cc:452-546, test_posterior_enum.py:300-508): greedily ADD
all rows; then sample_count times { sample_skip times { for each row: REMOVE,
ADD; [kind kernel]; [hyper kernel] }; record the latent state and the
engine-reported cross_cat.score_data() }.  The histogram of visited latents
must match exp(score) (multinomial goodness of fit > 5e-4 on the top-32
states).  Works against any Engine (CPU oracle or CUDA library).
"""
import itertools

import numpy as np
from scipy import stats

from loom_b200 import Model, RowBlock, sweep_actions
from loom_b200.synth import CLUSTERING, sample_py_partition, _sample_feature

TRUNCATE_COUNT = 32          # test_posterior_enum.py:50
MIN_GOODNESS_OF_FIT = 5e-4   # :51
SCORE_TOL = 1e-1             # :52
DENSITIES = [1.0, 0.5, 0.0]  # :57-61

LATENT_SIZES = [             # :65-79
    [1, 1, 2, 5, 15, 52, 203, 877, 4140, 21147, 115975, 678570, 4213597],
    [1, 1, 2, 5, 15, 52, 203, 877, 4140, 21147, 115975, 678570, 4213597],
    [1, 2, 6, 22, 94, 454, 2430, 14214, 89918, 610182, 4412798],
    [1, 5, 30, 205, 1555, 12880, 115155, 1101705],
    [1, 15, 240, 4065, 72465, 1353390],
    [1, 52, 2756, 148772, 8174244],
    [1, 203, 41412, 8489257],
    [1, 877, 770006],
    [1, 4140],
    [1, 21147],
    [1, 115975],
    [1, 678570],
    [1, 4213597],
]
BELL_NUMBERS = [1, 1, 2, 5, 15, 52, 203, 877, 4140, 21147, 115975, 678570, 4213597,
                27644437, 190899322, 1382958545, 10480142147, 82864869804, 682076806159]

PITMAN_YOR_GRID = [{"alpha": 2.0, "d": 0.1}, {"alpha": 10.0, "d": 0.1}]  # :85-88
HYPER_PRIOR = {              # :90-114
    "topology": PITMAN_YOR_GRID,
    "clustering": PITMAN_YOR_GRID,
    "bb": {"alpha": [0.5, 2.0], "beta": [0.5, 2.0]},
    "dd": {"alpha": [0.5, 1.5]},
    "gp": {"alpha": [0.5, 1.5], "inv_beta": [0.5, 1.5]},
    "nich": {"kappa": [0.5, 1.5], "mu": [-1.0, 1.0], "nu": [0.5, 1.5], "sigmasq": [0.5, 1.5]},
}

# stand-ins for distributions' EXAMPLES[0]['shared'] (not vendored)
EXAMPLE_SHARED = {
    "bb": dict(type="bb", alpha=0.5, beta=2.0),
    "dd": dict(type="dd", alphas=[0.5, 1.0, 2.0]),
    "dpd": dict(type="dpd", gamma=0.5, alpha=0.5, beta0=0.25, betas=[0.25, 0.25, 0.25]),
    "gp": dict(type="gp", alpha=1.0, inv_beta=1.0),
    "nich": dict(type="nich", mu=0.0, kappa=1.0, sigmasq=1.0, nu=1.0),
}
FEATURE_TYPES = list(EXAMPLE_SHARED)


def enum_partitions(count):
    if count == 0:
        yield ()
    elif count == 1:
        yield ((1,),)
    else:
        for p in enum_partitions(count - 1):
            yield p + ((count,),)
            for i, part in enumerate(p):
                yield p[:i] + (part + (count,),) + p[1 + i:]


def count_crosscats(rows, cols):
    return sum(BELL_NUMBERS[rows] ** len(kinds) for kinds in enum_partitions(cols))


def multinomial_goodness_of_fit(probs, counts, total_count, truncated=False):
    """distributions.util.multinomial_goodness_of_fit."""
    chi_squared, dof = 0.0, 0
    for p, c in zip(probs, counts):
        if p == 1:
            return 1.0 if c == total_count else 0.0
        if p > 0:
            mean = total_count * p
            variance = total_count * p * (1 - p)
            chi_squared += (c - mean) ** 2 / variance
            dof += 1
        elif c > 0:
            return 0.0
    if not truncated:
        dof -= 1
    if dof <= 0:
        return 1.0
    return float(stats.chi2.sf(chi_squared, dof))


def scores_to_probs(scores):
    scores = np.asarray(scores, np.float64)
    p = np.exp(scores - scores.max())
    return p / p.sum()


def generate_rows(object_count, feature_count, feature_type, density, rng):
    """test_posterior_enum.py:584-618: sample a table from the CrossCat prior."""
    shared = EXAMPLE_SHARED[feature_type]
    fa, kind_count = sample_py_partition(feature_count, CLUSTERING[0], CLUSTERING[1], rng)
    table = [[None] * feature_count for _ in range(object_count)]
    parts = [sample_py_partition(object_count, CLUSTERING[0], CLUSTERING[1], rng)
             for _ in range(kind_count)]
    for f, k in enumerate(fa):
        assign, m = parts[k]
        vals = _sample_feature(shared, assign, m, rng)
        for i in range(object_count):
            if rng.random() < density:
                v = vals[i]
                table[i][f] = float(np.uint32(v).view(np.float32)) if feature_type == "nich" else int(v)
    return table


def make_model(feature_count, feature_type):
    return Model([EXAMPLE_SHARED[feature_type]] * feature_count, [0] * feature_count,
                 kind_py=[CLUSTERING], topology=CLUSTERING)


def latent_of(engine):
    """parse_sample (test_posterior_enum.py:738-745) from the engine state."""
    n_kinds, f2k = engine.get_kinds()
    kinds = []
    for k in range(n_kinds):
        fids = frozenset(int(f) for f in np.nonzero(f2k == k)[0])
        if not fids:
            continue  # featureless kinds are skipped (src/loom.cc:530)
        a = engine.get_assignments(k)
        groups = {}
        for r, g in enumerate(a):
            groups.setdefault(int(g), []).append(r)
        kinds.append((fids, frozenset(frozenset(v) for v in groups.values())))
    return frozenset(kinds)


def run_posterior_enum(engine, model, rows, sample_count, sample_skip=10, infer_kinds=False,
                       hyper_prior=None, seed=0, empty_kind_count=32, iterations=32):
    engine.set_model(model)
    engine.set_rows(rows)
    if hyper_prior:
        engine.set_hyper_prior(hyper_prior)
    R = rows.n_rows
    counter = itertools.count(1)
    nseed = lambda: (seed << 20) + next(counter)
    engine.cat_sweep(sweep_actions(R), nseed())
    if infer_kinds:
        engine.init_featureless_kinds(empty_kind_count, nseed())
    refresh = sweep_actions(R, add=True, remove=True)
    fused = np.tile(refresh, sample_skip)
    samples = []
    for _ in range(sample_count):
        if not infer_kinds and not hyper_prior:
            engine.cat_sweep(fused, nseed())
        else:
            for _ in range(sample_skip):
                engine.cat_sweep(refresh, nseed())
                if infer_kinds:
                    engine.kind_proposer_score(fetch=False)
                    engine.kind_run(iterations, nseed())
                    engine.init_featureless_kinds(empty_kind_count, nseed())
                if hyper_prior:
                    engine.hyper_run(nseed())
        samples.append((latent_of(engine), engine.score_data()))
    return samples


def goodness_of_fit(samples, object_count, feature_count, infer_kinds, scores_override=None):
    """test_posterior_enum.py:433-495.  Returns (gof or None if inaccurate, info)."""
    counts_dict, scores_dict = {}, {}
    for latent, score in samples:
        if latent in counts_dict:
            counts_dict[latent] += 1
            # (the reference's own consistency assert is vacuous, :365-369; with
            # hyper inference the score of a latent legitimately varies)
            if scores_override is None:
                assert abs(score - scores_dict[latent]) < SCORE_TOL, "inconsistent score"
        else:
            counts_dict[latent] = 1
            scores_dict[latent] = score
    sample_count = len(samples)
    if scores_override is not None:
        latents = [lat for lat in scores_dict if lat in scores_override]
        scores_dict = scores_override
        sample_count = sum(counts_dict[lat] for lat in latents)
    else:
        latents = list(scores_dict)
    expected = count_crosscats(object_count, feature_count) if infer_kinds else BELL_NUMBERS[object_count]
    assert len(latents) <= expected, "programmer error"
    counts = np.array([counts_dict[k] for k in latents])
    probs = scores_to_probs([scores_dict[k] for k in latents])
    order = np.argsort(probs)[::-1][:TRUNCATE_COUNT]
    order = [i for i in order if sample_count * probs[i] * (1 - probs[i]) >= 1]
    truncated = len(order) < len(probs)
    if len(order) < 1:
        return None, dict(latents=len(latents), expected=expected)
    gof = multinomial_goodness_of_fit(probs[order], counts[order], sample_count, truncated)
    return gof, dict(latents=len(latents), expected=expected, n=sample_count)


def cat_dimensions(max_size):
    return [(r, c) for r, sizes in enumerate(LATENT_SIZES) for c, size in enumerate(sizes)
            if r > 1 and c > 0 and size <= max_size]


def kind_dimensions(max_size):
    return [(r, c) for r, sizes in enumerate(LATENT_SIZES) for c, size in enumerate(sizes)
            if r > 0 and c > 0 and size <= max_size and r + c > 2]
