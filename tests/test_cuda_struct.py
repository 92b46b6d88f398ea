"""GPU parity tests for the kind kernel (proposer accumulate + feature x kind
scoring + block Pitman-Yor + feature moves + ephemeral kinds) and the hyper
kernel, through the C ABI, against the float64 oracle."""
import numpy as np
import pytest

import pe_util
from loom_b200 import Engine, Model, sweep_actions
from loom_b200 import synth
from test_cuda_cat import both, mixed, assert_same_state, replay

pytestmark = pytest.mark.gpu


def lockstep_sweep(oracle_abi, o, g, acts, seed):
    picks, _ = g.cat_sweep(acts, seed=seed, debug=True)
    bad, _ = replay(oracle_abi, o, acts, seed, picks, None, 0)
    assert bad == 0


def test_feature_kind_scores_match_oracle(oracle_abi, cuda_abi):
    model, rows, _ = mixed(rows=2000, seed=4)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    lockstep_sweep(oracle_abi, o, g, sweep_actions(rows.n_rows), 5)
    for e in (o, g):
        e.init_featureless_kinds(6, seed=9)
    assert o.n_kinds() == g.n_kinds() == model.n_kinds + 6
    for k in range(o.n_kinds()):                      # same PY-prior partitions (same Philox draws)
        assert np.array_equal(o.get_assignments(k), g.get_assignments(k))
    Lo, Lg = o.kind_proposer_score(), g.kind_proposer_score()
    assert Lo.shape == Lg.shape == (model.n_features, model.n_kinds + 6)
    err = np.abs(Lg.astype(np.float64) - Lo) / np.maximum(1.0, np.abs(Lo))
    assert err.max() <= 1e-5, err.max()
    # the feature's own kind must reproduce the cross-cat statistics (validate_subset)
    fo, no = o.kind_run(8, seed=3)
    fg, ng = g.kind_run(8, seed=3)
    assert np.array_equal(fo, fg) and no == ng
    for k in range(o.n_kinds()):
        co, io, flo = o.get_groups(k)
        cg, ig, flg = g.get_groups(k)
        assert np.array_equal(co, cg) and np.array_equal(io, ig)       # moved integer statistics bit-exact
        assert np.allclose(flo, flg, rtol=1e-7, atol=1e-7)
    assert g.score_data() == pytest.approx(o.score_data(), rel=1e-6)
    # and the chain keeps running identically after the move
    for e in (o, g):
        e.init_featureless_kinds(4, seed=11)
    acts = sweep_actions(rows.n_rows, add=True, remove=True)
    lockstep_sweep(oracle_abi, o, g, acts, 12)
    for k in range(o.n_kinds()):
        assert np.array_equal(o.get_assignments(k), g.get_assignments(k))


def test_hyper_run_matches_oracle(oracle_abi, cuda_abi):
    model, rows, _ = synth.generate(1200, [("bb", 3), ("dd4", 2), ("dd16", 1), ("gp", 3), ("nich", 3)],
                                    density=0.7, seed=21)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    lockstep_sweep(oracle_abi, o, g, sweep_actions(rows.n_rows), 2)
    grids = {"topology": pe_util.PITMAN_YOR_GRID, "clustering": pe_util.PITMAN_YOR_GRID,
             "bb": {"alpha": [0.1, 0.5, 2.0, 8.0], "beta": [0.1, 0.5, 2.0, 8.0]},
             "dd": {"alpha": [0.1, 0.5, 1.5, 5.0]},
             "gp": {"alpha": [0.5, 1.5, 4.0], "inv_beta": [0.1, 0.5, 1.5]},
             "nich": {"mu": [-1.0, 0.0, 1.0], "kappa": [0.5, 1.5], "sigmasq": [0.5, 1.5, 4.0], "nu": [0.5, 1.5, 4.0]}}
    same = 0
    for seed in range(6):
        for e in (o, g):
            e.set_hyper_prior(grids)
            e.hyper_run(seed)
        so, ao, ko, to = o.get_hypers()
        sg, ag, kg, tg = g.get_hypers()
        hp_o = np.array([[s.hp[i] for i in range(4)] for s in so])
        hp_g = np.array([[s.hp[i] for i in range(4)] for s in sg])
        assert np.array_equal(ko, kg) and np.array_equal(to, tg)     # host float64 tasks: identical
        if np.array_equal(hp_o, hp_g) and np.array_equal(ao, ag):
            same += 1
        else:   # a near-tie between grid points may flip one draw: resynchronise and go on
            n_diff = int((hp_o != hp_g).sum() + (ao != ag).sum())
            assert n_diff <= 2, n_diff
            break
        assert g.score_data() == pytest.approx(o.score_data(), rel=1e-6)
        acts = sweep_actions(rows.n_rows, add=True, remove=True)
        lockstep_sweep(oracle_abi, o, g, acts, 100 + seed)
    assert same >= 3


@pytest.mark.parametrize("feature_type", ["bb", "dd", "gp", "nich"])
def test_posterior_enum_kinds_on_gpu(cuda_abi, feature_type):
    """infer_kinds (test_posterior_enum.py:164-188) on the CUDA kernels."""
    rng = np.random.default_rng(99)
    for density in (1.0, 0.5):
        for (R, Cn) in [(1, 3), (2, 2), (3, 2)]:
            model = pe_util.make_model(Cn, feature_type)
            rows = pe_util.RowBlock.from_table(model, pe_util.generate_rows(R, Cn, feature_type, density, rng))
            n = 10 * pe_util.LATENT_SIZES[R][Cn]
            gof, info = pe_util.goodness_of_fit(
                pe_util.run_posterior_enum(Engine(cuda_abi), model, rows, n, infer_kinds=True, seed=R * 10 + Cn,
                                           empty_kind_count=8), R, Cn, True)
            if gof is not None and gof <= pe_util.MIN_GOODNESS_OF_FIT:
                gof, info = pe_util.goodness_of_fit(
                    pe_util.run_posterior_enum(Engine(cuda_abi), model, rows, 4 * n, infer_kinds=True,
                                               seed=555 + R * 10 + Cn, empty_kind_count=8), R, Cn, True)
            assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (feature_type, density, R, Cn, gof, info)


def test_posterior_enum_hypers_on_gpu(cuda_abi, oracle_abi):
    """infer_feature_hypers (test_posterior_enum.py:192-223) for one BB and one NICH grid on the GPU;
    expected probabilities come from fixed-hyper runs on the oracle."""
    from test_oracle_structure import _fixed_scores
    for feature_type, param in [("bb", "alpha"), ("nich", "sigmasq")]:
        grid = pe_util.HYPER_PRIOR[feature_type][param]
        R = 3
        rng = np.random.default_rng(7)
        base = pe_util.make_model(1, feature_type)
        rows = pe_util.RowBlock.from_table(base, pe_util.generate_rows(R, 1, feature_type, 1.0, rng))
        variants = []
        for v in grid:
            f = dict(base.features[0])
            f[param] = v
            variants.append(Model([f], [0], kind_py=[pe_util.CLUSTERING], topology=pe_util.CLUSTERING))
        fixed = _fixed_scores(oracle_abi, base, rows, variants, R, 1, False, seed=R)
        samples = pe_util.run_posterior_enum(Engine(cuda_abi), base, rows, 40 * pe_util.LATENT_SIZES[R][1],
                                             hyper_prior={feature_type: {param: grid}}, seed=70 + R)
        gof, info = pe_util.goodness_of_fit(samples, R, 1, False, scores_override=fixed)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (feature_type, param, gof, info)


def test_ephemeral_kinds_seated_ahead_of_time_equal_the_direct_draw(oracle_abi, cuda_abi):
    """lb_prepare_featureless_kinds draws the next init_featureless_kinds on background threads while the device
    sweeps; the kinds it yields are the kinds the direct call yields (and the oracle's), and a draw prepared for another
    situation (other seed) is discarded"""
    model, rows, _ = mixed(rows=1500, seed=6)
    o, a = both(oracle_abi, cuda_abi, model, rows)
    b = Engine(cuda_abi)
    b.set_model(model)
    b.set_rows(rows)
    first = sweep_actions(rows.n_rows)
    again = sweep_actions(rows.n_rows, add=True, remove=True)
    lockstep_sweep(oracle_abi, o, a, first, 5)
    b.cat_sweep(first, seed=5)
    a.prepare_featureless_kinds(5, seed=9)           # drawn under the next sweep ...
    lockstep_sweep(oracle_abi, o, a, again, 6)
    b.cat_sweep(again, seed=6)
    b.prepare_featureless_kinds(5, seed=1234)        # ... and one prepared for a different seed: must not be used
    for e in (o, a, b):
        e.init_featureless_kinds(5, seed=9)
    for k in range(o.n_kinds()):
        assert np.array_equal(o.get_assignments(k), a.get_assignments(k)), k
        assert np.array_equal(a.get_assignments(k), b.get_assignments(k)), k
        assert np.array_equal(o.get_groups(k)[0], a.get_groups(k)[0])
    assert a.kind_info(o.n_kinds() - 1).sample_size == rows.n_rows


def test_featureless_kinds_with_thousands_of_groups_replay(oracle_abi, cuda_abi):
    """K3f (lb_sweep_flat.cuh): ephemeral kinds whose clustering draw gives thousands of groups (loom's grid reaches
    alpha = 50, d = 0.36; here alpha = 300, d = 0.5 on 12 000 rows -> ~3 000 groups, 256 per lane block, so both levels
    of the warp scan and the table growth path run).  Every pick replays on the oracle, final state bit-exact."""
    model, rows, _ = synth.generate(12000, [("bb", 3)], density=0.8, seed=2, single_kind=True)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    grids = {"clustering": [(300.0, 0.5)]}
    for e in (o, g):
        e.set_hyper_prior(grids)
    lockstep_sweep(oracle_abi, o, g, sweep_actions(rows.n_rows), 3)
    for e in (o, g):
        e.init_featureless_kinds(2, seed=4)             # seated from PY(300, 0.5): thousands of groups each
    n_groups = [g.kind_info(k).n_groups for k in range(g.n_kinds())]
    assert max(n_groups[1:]) > 1500, n_groups
    acts = sweep_actions(rows.n_rows, add=True, remove=True)
    lockstep_sweep(oracle_abi, o, g, acts, 5)
    for k in range(o.n_kinds()):
        assert np.array_equal(o.get_assignments(k), g.get_assignments(k)), k
        co, cg = o.get_groups(k)[0], g.get_groups(k)[0]
        assert np.array_equal(co, cg), k
    # and once more from the state the device left (counts / id maps written back by the flat kernel)
    lockstep_sweep(oracle_abi, o, g, acts, 6)
    for k in range(o.n_kinds()):
        assert np.array_equal(o.get_assignments(k), g.get_assignments(k)), k
