"""GPU parity tests for the cat-kernel path, called through the C ABI
(libloomb200.so) and checked against the float64 CPU oracle.

Bars (BASELINE.json north_star): integer sufficient statistics bit-exact for
the same assignments; per-row x per-group scores within 1e-5 relative
(|gpu - oracle| <= 1e-5 * max(1, |oracle|)); sampled assignments consistent
with the shared Philox uniforms and pass loom's posterior-enum chi-square.
"""
import ctypes as C

import numpy as np
import pytest

import pe_util
from loom_b200 import cat_sweep_multi, Engine, Model, sweep_actions, UNASSIGNED
from loom_b200 import synth

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5


def rel_err(gpu, ref):
    return np.abs(gpu.astype(np.float64) - ref) / np.maximum(1.0, np.abs(ref))


def mixed(rows=1500, seed=11, alpha=None):
    model, rb, truth = synth.generate(rows, [("bb", 5), ("dd4", 3), ("dd16", 3), ("dd256", 2),
                                             ("dpd7", 2), ("gp", 4), ("nich", 4)],
                                      density=0.6, seed=seed)
    if alpha is not None:
        model.kind_py[:, 0] = alpha
    return model, rb, truth


def both(oracle_abi, cuda_abi, model, rows):
    o, g = Engine(oracle_abi), Engine(cuda_abi)
    for e in (o, g):
        e.set_model(model)
        e.set_rows(rows)
    return o, g


def replay(oracle_abi, o, actions, seed, picks, scores, stride, offset=0, tol=2e-5):
    lib = oracle_abi.lib
    lib.lo_cat_replay.restype = C.c_int64
    lib.lo_cat_replay.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64,
                                  C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                  C.c_int32, C.c_double, C.POINTER(C.c_double)]
    err = C.c_double()
    actions = np.ascontiguousarray(actions, np.uint32)
    bad = lib.lo_cat_replay(o.h, actions.ctypes.data_as(C.POINTER(C.c_uint32)), len(actions), seed,
                            offset, picks.ctypes.data_as(C.POINTER(C.c_uint32)),
                            scores.ctypes.data_as(C.POINTER(C.c_float)) if scores is not None else None,
                            stride, tol, C.byref(err))
    assert bad >= 0, oracle_abi.last_error()
    return bad, err.value


def assert_same_state(o, g, model, flt_rtol=1e-7):
    for k in range(model.n_kinds):
        co, io, fo = o.get_groups(k)
        cg, ig, fg = g.get_groups(k)
        assert np.array_equal(co, cg), k
        assert np.array_equal(io, ig), k           # bit-exact integer statistics
        assert np.allclose(fo, fg, rtol=flt_rtol, atol=1e-7), k
        assert np.array_equal(o.get_assignments(k), g.get_assignments(k)), k


@pytest.mark.parametrize("kernel", ["spec", "classic"])
def test_sweep_matches_oracle_by_replay(oracle_abi, cuda_abi, monkeypatch, kernel):
    """both sweep kernels: K3s (scores gathered one ADD ahead, touched column recomputed) and the classic form"""
    monkeypatch.setenv("LB_SWEEP_KERNEL", kernel)
    model, rows, _ = mixed()
    o, g = both(oracle_abi, cuda_abi, model, rows)
    R = rows.n_rows
    acts = np.concatenate([sweep_actions(R), sweep_actions(R, add=True, remove=True)])
    stride = 128
    picks, scores = g.cat_sweep(acts, seed=42, debug=True, debug_stride=stride)
    bad, err = replay(oracle_abi, o, acts, 42, picks, scores, stride)
    assert err <= SCORE_TOL, err
    assert bad == 0
    assert_same_state(o, g, model)
    assert g.score_data() == pytest.approx(o.score_data(), rel=1e-6)


@pytest.mark.parametrize("types,rows_n", [
    ([("bb", 1)], 5),                                                   # the smallest posterior-enum case
    ([("bb", 5), ("dd16", 4), ("dpd7", 2)], 700),                       # narrow kinds: one warp per (chain, kind)
    ([("bb", 30), ("dd16", 40), ("dd256", 10)], 1200),                  # 8- and 16-warp CTAs
])
def test_sweep_categorical_only_instantiation_replays(oracle_abi, cuda_abi, monkeypatch, types, rows_n):
    """K3s compiled without the GP / NICH paths (launches whose kinds hold Bernoulli / categorical columns only, C3):
    every pick replays on the oracle, final state bit-exact; and it equals the general instantiation's sweep"""
    monkeypatch.setenv("LB_SWEEP_KERNEL", "spec")
    model, rows, _ = synth.generate(rows_n, types, density=0.7, seed=9, single_kind=len(types) == 1)
    acts = np.concatenate([sweep_actions(rows_n)] + [sweep_actions(rows_n, add=True, remove=True)] * 3)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    picks, scores = g.cat_sweep(acts, seed=11, debug=True, debug_stride=64)
    bad, err = replay(oracle_abi, o, acts, 11, picks, scores, 64)
    assert bad == 0 and err <= SCORE_TOL, (bad, err)
    assert_same_state(o, g, model)
    monkeypatch.setenv("LB_SPEC_CATONLY", "0")
    g2 = Engine(cuda_abi)
    g2.set_model(model)
    g2.set_rows(rows)
    picks2, _ = g2.cat_sweep(acts, seed=11, debug=True)
    assert np.array_equal(picks, picks2)


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_sweep_every_cta_width_and_multi_chain(oracle_abi, cuda_abi, monkeypatch, warps):
    """1/2/4/8 warps per (chain, kind): each width replays exactly on the oracle, and the
    many-chains launch (lb_cat_sweep_multi, rows shared by lb_share_rows) leaves every chain in
    the state its own single-chain sweep reaches."""
    monkeypatch.setenv("LB_SWEEP_WARPS", str(warps))
    model, rows, _ = mixed(rows=900, seed=5)
    R = rows.n_rows
    acts = np.concatenate([sweep_actions(R), sweep_actions(R, add=True, remove=True)])
    seeds = [21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31]     # 11 chains: a full block of 8 and a ragged one
    singles = []
    for c, seed in enumerate(seeds):
        g = Engine(cuda_abi)
        g.set_model(model)
        g.set_rows(rows)
        picks, scores = g.cat_sweep(acts, seed=seed, debug=True, debug_stride=64)
        if c < 2:
            o = Engine(oracle_abi)
            o.set_model(model)
            o.set_rows(rows)
            bad, err = replay(oracle_abi, o, acts, seed, picks, scores, 64)
            assert bad == 0 and err <= SCORE_TOL
            assert_same_state(o, g, model)
        singles.append(g)
    owner = Engine(cuda_abi)
    owner.set_model(model)
    owner.set_rows(rows)
    chains = [owner]
    for _ in seeds[1:]:
        g = Engine(cuda_abi)
        g.set_model(model)
        g.share_rows(owner)
        chains.append(g)
    cat_sweep_multi(chains, acts, seeds)
    for a, b in zip(singles, chains):
        for k in range(model.n_kinds):
            assert np.array_equal(a.get_assignments(k), b.get_assignments(k))
            assert all(np.array_equal(x, y) for x, y in zip(a.get_groups(k), b.get_groups(k)))
    # chains differ from each other (independent seeds)
    assert not np.array_equal(chains[0].get_assignments(0), chains[1].get_assignments(0))
    # a second call continues from the current state; errors name the chain
    cat_sweep_multi(chains, sweep_actions(R, add=False), seeds)
    assert all(c.kind_info(0).sample_size == 0 for c in chains)
    with pytest.raises(Exception, match="remove of unassigned row"):
        cat_sweep_multi(chains, sweep_actions(4, add=False), seeds)


def test_sweep_grows_tables_and_renumbers_ids(oracle_abi, cuda_abi, monkeypatch):
    # many groups: exercises Gcap growth (64 -> 128 -> ...), the >128-group
    # chunked scorer and the global-id renumbering (> 4096 group births)
    monkeypatch.setenv("LB_INITIAL_GCAP", "64")
    model, rows, _ = mixed(rows=2500, seed=3, alpha=60.0)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    R = rows.n_rows
    acts = np.concatenate([sweep_actions(R)] + [sweep_actions(R, add=True, remove=True)] * 12)
    picks, _ = g.cat_sweep(acts, seed=7, debug=True)
    assert max(g.kind_info(k).n_groups for k in range(model.n_kinds)) > 128
    bad, _ = replay(oracle_abi, o, acts, 7, picks, None, 0)
    assert bad == 0
    assert_same_state(o, g, model, flt_rtol=1e-6)


@pytest.mark.parametrize("warps", ["spec", 1, 8])
def test_multi_chain_sweep_grows_tables(cuda_abi, monkeypatch, warps):
    """Table growth and id renumbering inside the many-chain launch: chains stop at different
    actions, the host grows their tables and relaunches only the unfinished (chain, kind) pairs;
    every chain must end where its own single-chain sweep ends."""
    monkeypatch.setenv("LB_INITIAL_GCAP", "64")
    if warps == "spec":
        monkeypatch.setenv("LB_SWEEP_KERNEL", "spec")          # K3s: wide and narrow CTAs on two streams
    else:
        monkeypatch.setenv("LB_SWEEP_WARPS", str(warps))       # the classic kernel at that CTA width
    model, rows, _ = mixed(rows=1200, seed=4, alpha=60.0)
    R = rows.n_rows
    acts = np.concatenate([sweep_actions(R)] + [sweep_actions(R, add=True, remove=True)] * 6)
    seeds = [3, 4, 5, 6, 7]
    singles = []
    for seed in seeds:
        g = Engine(cuda_abi)
        g.set_model(model)
        g.set_rows(rows)
        g.cat_sweep(acts, seed=seed)
        singles.append(g)
    assert max(singles[0].kind_info(k).n_groups for k in range(model.n_kinds)) > 64   # growth happened
    owner = Engine(cuda_abi)
    owner.set_model(model)
    owner.set_rows(rows)
    chains = [owner]
    for _ in seeds[1:]:
        g = Engine(cuda_abi)
        g.set_model(model)
        g.share_rows(owner)
        chains.append(g)
    cat_sweep_multi(chains, acts, seeds)
    for a, b in zip(singles, chains):
        for k in range(model.n_kinds):
            assert np.array_equal(a.get_assignments(k), b.get_assignments(k))
            assert all(np.array_equal(x, y) for x, y in zip(a.get_groups(k), b.get_groups(k)))


def test_remove_everything_returns_to_empty(cuda_abi):
    model, rows, _ = mixed(rows=400)
    g = Engine(cuda_abi)
    g.set_model(model)
    g.set_rows(rows)
    g.cat_sweep(sweep_actions(rows.n_rows), seed=1)
    g.cat_sweep(sweep_actions(rows.n_rows, add=False), seed=2)
    for k in range(model.n_kinds):
        counts, ints, flts = g.get_groups(k)
        assert len(counts) == 1 and counts.sum() == 0 and not ints.any()
        assert np.allclose(flts, 0, atol=1e-9)
        assert (g.get_assignments(k) == UNASSIGNED).all()


def test_errors_are_loud(cuda_abi):
    model, rows, _ = mixed(rows=64)
    g = Engine(cuda_abi)
    g.set_model(model)
    g.set_rows(rows)
    g.cat_sweep(sweep_actions(10), seed=1)
    with pytest.raises(Exception, match="duplicate row"):
        g.cat_sweep(sweep_actions(1), seed=1)
    g2 = Engine(cuda_abi)
    g2.set_model(model)
    g2.set_rows(rows)
    with pytest.raises(Exception, match="unassigned"):
        g2.cat_sweep(sweep_actions(1, add=False), seed=1)
    bad = synth.generate(64, [("dd4", 2)], seed=1)
    bad[1].cat8[0, 3] = 9
    bad[1].mask[:] = 0xFFFFFFFF
    g3 = Engine(cuda_abi)
    g3.set_model(bad[0])
    with pytest.raises(Exception, match="out of range"):
        g3.set_rows(bad[1])


def _load_state(src, dst, model):
    uniqs = []
    for k in range(model.n_kinds):
        a = src.get_assignments(k)
        counts, ints, flts = src.get_groups(k)
        keep = np.nonzero(counts)[0]
        remap = np.full(len(counts) + 1, UNASSIGNED, np.uint32)
        remap[keep] = np.arange(len(keep), dtype=np.uint32)
        dst.set_groups(k, len(keep), counts[keep], ints[:, keep], flts[:, keep])
        dst.set_assignments(k, np.where(a == UNASSIGNED, UNASSIGNED, remap[np.minimum(a, len(counts))]))
        uniqs.append(keep)
    return uniqs


@pytest.mark.parametrize("seed", [0, 1])
def test_score_rows_parity(oracle_abi, cuda_abi, seed):
    model, rows, _ = mixed(rows=3000, seed=20 + seed)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    o.cat_sweep(sweep_actions(2000), seed=5)
    keeps = _load_state(o, g, model)
    worst = 0.0
    for k in range(model.n_kinds):
        so = o.score_rows(k).astype(np.float64)
        sg = g.score_rows(k)
        so = np.concatenate([so[:, keeps[k]], so[:, [c for c in range(so.shape[1]) if c not in set(keeps[k])]]], axis=1)
        assert so.shape == sg.shape
        worst = max(worst, rel_err(sg, so).max())
    assert worst <= SCORE_TOL, worst


@pytest.mark.parametrize("types,rows_n,planes", [
    ([("bb", 40)], 1000, 3),                                  # Bernoulli block: 80 operand columns, 2 K chunks
    ([("bb", 7), ("dd16", 9), ("dpd5", 3)], 2777, 3),         # ragged tile, DPD ids from the wide columns
    ([("dd16", 24), ("dd256", 3)], 1500, 3),                  # a DD-256 column spans 4 chunks of its own
    ([("bb", 12), ("dd16", 12)], 1300, 2),                    # two planes (16 mantissa bits) still inside 1e-5
])
def test_score_rows_tensor_core_matches_oracle(oracle_abi, cuda_abi, monkeypatch, types, rows_n, planes):
    """K1t (lb_score_tc.cu): one-hot x score-table contraction on tcgen05 with split precision == the float64
    oracle within 1e-5 * max(1, |score|), and == the gather kernel to fp32 summation order"""
    model, rows, _ = synth.generate(rows_n, types, density=0.6, seed=31, single_kind=len(types) == 1)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    o.cat_sweep(sweep_actions(rows_n), seed=5)                # ~dozens of groups per kind
    keeps = _load_state(o, g, model)
    monkeypatch.setenv("LB_K1_TC_PLANES", str(planes))
    for k in range(model.n_kinds):
        so = o.score_rows(k).astype(np.float64)
        so = np.concatenate([so[:, keeps[k]], so[:, [c for c in range(so.shape[1]) if c not in set(keeps[k])]]], axis=1)
        monkeypatch.setenv("LB_K1_TC", "0")
        gather = g.score_rows(k)
        n0 = g.launch_count()[2]
        monkeypatch.setenv("LB_K1_TC", "1")
        tc = g.score_rows(k)
        assert g.launch_count()[2] - n0 == 3                  # score table + operand table + ONE contraction launch
        assert tc.shape == so.shape
        assert rel_err(tc, so).max() <= SCORE_TOL, (k, rel_err(tc, so).max())
        assert np.abs(tc.astype(np.float64) - gather).max() <= (2e-5 if planes == 3 else 2e-4) * max(1.0, np.abs(gather).max())
        part = g.score_rows(k, 130, 391)                      # a row range that starts inside a tile
        assert np.array_equal(part, tc[130:391])


def test_score_rows_large_groups_float_accuracy(oracle_abi, cuda_abi):
    # groups holding ~1e5 rows: fp32 lgamma/log cancellation must not leak in
    model = Model([dict(type="gp", alpha=2.0, inv_beta=0.5), dict(type="nich", mu=0.0, kappa=1.0, sigmasq=1.0, nu=2.0),
                   dict(type="bb", alpha=0.5, beta=0.5), dict(type="dd", alphas=[0.5] * 8)], [0, 0, 0, 0])
    order = np.argsort([{"bb": 0, "dd": 1, "gp": 3, "nich": 4}[f["type"]] for f in model.features], kind="stable")
    model = Model([model.features[i] for i in order], [0] * 4)
    rng = np.random.default_rng(0)
    R = 64
    table = [[int(rng.integers(2)), int(rng.integers(8)), int(rng.poisson(30.0 if r % 2 else 3.0)),
              float(rng.normal(0, 3))] for r in range(R)]
    rows = pe_util.RowBlock.from_table(model, table)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    counts = np.uint32([120000, 35, 7, 2500000])
    n = counts.astype(np.uint64)
    ints = np.zeros((2 + 9 + 2 + 1, 4), np.uint32)
    ints[0] = n * 3 // 10; ints[1] = n - ints[0]                     # bb heads/tails
    ints[2:10] = (n // 8)[None, :]; ints[2] += (n - 8 * (n // 8)).astype(np.uint32)  # dd counts
    ints[11] = n; ints[12] = (n * np.uint64([3, 30, 4, 12])).astype(np.uint32)      # gp count,sum
    ints[13] = n                                                     # nich count
    flts = np.zeros((3, 4))
    flts[0] = n * 2.0                                                # gp log_prod (unused by predictive)
    flts[1] = [0.25, -3.0, 1.0, 10.0]                                # nich mean
    flts[2] = n * np.array([1.0, 4.0, 0.5, 9.0])                     # nich ctv
    for e in (o, g):
        e.set_groups(0, 4, counts, ints, flts)
    so, sg = o.score_rows(0).astype(np.float64), g.score_rows(0)
    assert rel_err(sg, so).max() <= SCORE_TOL, rel_err(sg, so).max()


def test_query_score_parity(oracle_abi, cuda_abi):
    """QueryServer Score for one sample (lb_query_score): sum over kinds of log-sum-exp of
    score_value, within the score tolerance of the float64 oracle."""
    model, rows, truth = mixed(rows=3000, seed=21)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    for k in range(model.n_kinds):
        uniq, inv = np.unique(truth["assign"][k], return_inverse=True)
        for e in (o, g):
            e.set_groups(k, len(uniq))
            e.set_assignments(k, inv.astype(np.uint32))
    for e in (o, g):
        e.accumulate_rows()
    so, sg = o.query_score().astype(np.float64), g.query_score()
    assert rel_err(sg, so).max() <= SCORE_TOL, rel_err(sg, so).max()
    assert np.array_equal(g.query_score(100, 164), sg[100:164])     # sub-ranges address the same rows


@pytest.mark.parametrize("n_rows,lo,hi", [(4000, 0, 2000), (4003, 3, 1999), (4002, 256, 4002)])
def test_accumulate_parity(oracle_abi, cuda_abi, n_rows, lo, hi):
    # (4000, 0, 2000) takes the four-rows-per-load path, the others the unaligned one
    model, rows, truth = mixed(rows=n_rows, seed=8)
    o, g = both(oracle_abi, cuda_abi, model, rows)
    for k in range(model.n_kinds):
        a = truth["assign"][k].astype(np.uint32)
        a[::7] = UNASSIGNED                       # ragged: some rows unassigned
        uniq, inv = np.unique(a[a != UNASSIGNED], return_inverse=True)
        packed = np.full(len(a), UNASSIGNED, np.uint32)
        packed[a != UNASSIGNED] = inv
        for e in (o, g):
            e.set_groups(k, len(uniq))
            e.set_assignments(k, packed)
    for e in (o, g):
        e.accumulate_rows()
    assert_same_state(o, g, model, flt_rtol=1e-9)
    # removing part of the rows again stays exact on the integer statistics
    for e in (o, g):
        e.accumulate_rows(row_begin=lo, row_end=hi, sign=-1)
    for k in range(model.n_kinds):
        co, io, fo = o.get_groups(k)
        cg, ig, fg = g.get_groups(k)
        assert np.array_equal(co, cg) and np.array_equal(io, ig)
        assert np.allclose(fo, fg, rtol=1e-6, atol=1e-6)


def test_empty_and_ragged_inputs(oracle_abi, cuda_abi):
    # density 0 (no observed cells), 1 row, R not a multiple of 32
    for rows_n, dens in [(1, 1.0), (33, 0.0), (95, 0.3)]:
        model, rows, _ = synth.generate(rows_n, [("bb", 2), ("dd4", 1), ("gp", 1), ("nich", 1)],
                                        density=dens, seed=rows_n)
        o, g = both(oracle_abi, cuda_abi, model, rows)
        acts = np.concatenate([sweep_actions(rows_n), sweep_actions(rows_n, add=True, remove=True)])
        picks, scores = g.cat_sweep(acts, seed=3, debug=True, debug_stride=64)
        bad, err = replay(oracle_abi, o, acts, 3, picks, scores, 64)
        assert bad == 0 and err <= SCORE_TOL
        assert_same_state(o, g, model)
    g = Engine(cuda_abi)
    g.set_model(model)
    g.set_rows(rows)
    g.cat_sweep(np.zeros(0, np.uint32), seed=1)   # empty action list is a no-op


@pytest.mark.parametrize("feature_type", pe_util.FEATURE_TYPES)
def test_posterior_enum_cats_on_gpu(cuda_abi, feature_type):
    """loom's own chi-square test (test_posterior_enum.py:139-160) on the CUDA kernel."""
    rng = np.random.default_rng(321)
    for density in (1.0, 0.5):
        for (R, Cn) in [(2, 2), (3, 1), (3, 2), (4, 1), (5, 1)]:
            model = pe_util.make_model(Cn, feature_type)
            rows = pe_util.RowBlock.from_table(model, pe_util.generate_rows(R, Cn, feature_type, density, rng))
            n = 10 * pe_util.LATENT_SIZES[R][1]
            gof, info = pe_util.goodness_of_fit(
                pe_util.run_posterior_enum(Engine(cuda_abi), model, rows, n, seed=R * 10 + Cn), R, Cn, False)
            if gof is not None and gof <= pe_util.MIN_GOODNESS_OF_FIT:
                # two-stage test: a 5e-4 false alarm must not repeat on 4x the samples
                gof, info = pe_util.goodness_of_fit(
                    pe_util.run_posterior_enum(Engine(cuda_abi), model, rows, 4 * n, seed=977 + R * 10 + Cn), R, Cn, False)
            assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (feature_type, density, R, Cn, gof, info)
