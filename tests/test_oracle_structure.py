"""Oracle self-checks for the kind kernel and hyper kernel: loom's
posterior-enumeration tests infer_kinds / infer_*_hypers
(loom/test/test_posterior_enum.py:164-280) run on the oracle."""
import numpy as np
import pytest

import pe_util
from loom_b200 import Engine, Model


def _run_case(abi, R, C, feature_type, density, infer_kinds, hyper=None, seed=0, max_kind=50):
    rng = np.random.default_rng(123 + seed)
    model = pe_util.make_model(C, feature_type)
    rows = pe_util.RowBlock.from_table(model, pe_util.generate_rows(R, C, feature_type, density, rng))
    size = pe_util.LATENT_SIZES[R][C] if infer_kinds else pe_util.LATENT_SIZES[R][1]
    e = Engine(abi)
    samples = pe_util.run_posterior_enum(e, model, rows, 10 * size, infer_kinds=infer_kinds,
                                         hyper_prior=hyper, seed=seed)
    return model, rows, samples


@pytest.mark.parametrize("feature_type", pe_util.FEATURE_TYPES)
@pytest.mark.parametrize("density", pe_util.DENSITIES)
def test_posterior_enum_kinds(oracle_abi, feature_type, density):
    for i, (R, C) in enumerate(pe_util.kind_dimensions(50)):
        _, _, samples = _run_case(oracle_abi, R, C, feature_type, density, True, seed=i)
        gof, info = pe_util.goodness_of_fit(samples, R, C, infer_kinds=True)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)


def _fixed_scores(abi, model, rows, variants, R, C, infer_kinds, seed):
    """process_fixed_samples (test_posterior_enum.py:398-416): run once per
    fixed hyper setting, log-sum-exp the scores of latents seen in all runs."""
    per = []
    size = pe_util.LATENT_SIZES[R][C] if infer_kinds else pe_util.LATENT_SIZES[R][1]
    for i, m in enumerate(variants):
        e = Engine(abi)
        s = pe_util.run_posterior_enum(e, m, rows, 10 * size, infer_kinds=infer_kinds,
                                       seed=1000 * seed + i)
        d = {}
        for lat, sc in s:
            d[lat] = sc
        per.append(d)
    common = set.intersection(*[set(d) for d in per])
    return {lat: float(np.logaddexp.reduce([d[lat] for d in per])) for lat in common}


HYPER_CASES = [(t, p) for t in ("bb", "dd", "gp", "nich") for p in pe_util.HYPER_PRIOR[t]]


@pytest.mark.parametrize("feature_type,param", HYPER_CASES)
def test_posterior_enum_feature_hypers(oracle_abi, feature_type, param):
    """infer_feature_hypers (test_posterior_enum.py:192-223), densities 1 and .5."""
    grid = pe_util.HYPER_PRIOR[feature_type][param]
    for density in (1.0, 0.5):
        for R in (2, 3, 4):
            rng = np.random.default_rng(77 + R)
            base = pe_util.make_model(1, feature_type)
            rows = pe_util.RowBlock.from_table(
                base, pe_util.generate_rows(R, 1, feature_type, density, rng))
            variants = []
            if feature_type == "dd":
                import itertools
                dim = len(base.features[0]["alphas"])
                for combo in itertools.product(grid, repeat=dim):
                    f = dict(base.features[0], alphas=list(combo))
                    variants.append(Model([f], [0], kind_py=[pe_util.CLUSTERING], topology=pe_util.CLUSTERING))
            else:
                for v in grid:
                    f = dict(base.features[0])
                    f[param] = v
                    variants.append(Model([f], [0], kind_py=[pe_util.CLUSTERING], topology=pe_util.CLUSTERING))
            fixed = _fixed_scores(oracle_abi, base, rows, variants, R, 1, False, seed=R)
            e = Engine(oracle_abi)
            samples = pe_util.run_posterior_enum(
                e, base, rows, 10 * pe_util.LATENT_SIZES[R][1],
                hyper_prior={feature_type: {param: grid}}, seed=50 + R)
            gof, info = pe_util.goodness_of_fit(samples, R, 1, False, scores_override=fixed)
            assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (feature_type, param, R, gof, info)


@pytest.mark.parametrize("which", ["clustering", "topology"])
def test_posterior_enum_py_hypers(oracle_abi, which):
    """infer_clustering_hypers / infer_topology_hypers (test_posterior_enum.py:227-280)."""
    grid = pe_util.PITMAN_YOR_GRID
    infer_kinds = which == "topology"
    dims = [(2, 2), (3, 2)] if infer_kinds else [(2, 1), (3, 1), (4, 1)]
    for (R, C) in dims:
        rng = np.random.default_rng(5 + R)
        base = pe_util.make_model(C, "bb")
        rows = pe_util.RowBlock.from_table(base, pe_util.generate_rows(R, C, "bb", 1.0, rng))
        variants = []
        for p in grid:
            py = (p["alpha"], p["d"])
            if which == "clustering":
                variants.append(Model(base.features, base.f2k, kind_py=[py], topology=pe_util.CLUSTERING))
            else:
                variants.append(Model(base.features, base.f2k, kind_py=[pe_util.CLUSTERING], topology=py))
        fixed = _fixed_scores(oracle_abi, base, rows, variants, R, C, infer_kinds, seed=R)
        e = Engine(oracle_abi)
        size = pe_util.LATENT_SIZES[R][C] if infer_kinds else pe_util.LATENT_SIZES[R][1]
        samples = pe_util.run_posterior_enum(e, base, rows, 10 * size, infer_kinds=infer_kinds,
                                             hyper_prior={which: grid}, seed=60 + R)
        gof, info = pe_util.goodness_of_fit(samples, R, C, infer_kinds, scores_override=fixed)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (which, R, C, gof, info)
