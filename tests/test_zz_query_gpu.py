"""GPU runs of the rows SURVEY.md 8f added last: the query path (n3) -- lb_query_sample against the oracle draw for draw, the
product binary loom_query on the device against scores recomputed by the oracle -- and the tare / sparsify passes (n4).  Kept in the LAST test file so that a
failure here cannot hide the results of the earlier GPU tests under `pytest -x`."""
import os
import subprocess

import numpy as np
import pytest

from loom_b200 import _abi
from test_pbio import io_lib  # noqa: F401  (autouse fixture builds the codec)
from test_z_loom_infer import HOST, infer_on_oracle  # noqa: F401
import test_query_server as qs


# ------------------------------------------------------------------------------------------ query sampler on the device
@pytest.mark.gpu
def test_query_sample_matches_oracle_on_the_gpu(oracle_abi, cuda_abi):
    """lb_query_sample draw for draw against the oracle (same Philox counters, float64 samplers on both
    sides; the group pick inverts fp32 scores on the device and float64 ones on the oracle, so a draw
    that lands within rounding of a CDF step may differ), then the device's own draws against the
    device's own scores."""
    import test_query as tq
    model, o = tq.trained(oracle_abi)
    observe = [f for f in range(model.n_features) if f % 3 == 0]
    cond = tq.conditioning_row(model, observe)
    to_sample = [f for f in range(model.n_features) if f % 3 != 0]
    qo = tq.frozen(oracle_abi, model, o, [cond])
    qg = tq.frozen(cuda_abi, model, o, [cond])
    n = tq.N_DRAWS
    vo, go = qo.query_sample(0, to_sample, n, seed=7, groups=True)
    vg, gg = qg.query_sample(0, to_sample, n, seed=7, groups=True)
    assert np.array_equal(go == _abi.UNASSIGNED, gg == _abi.UNASSIGNED)
    assert (go == gg).mean() > 0.9995
    for j, f in enumerate(to_sample):
        same_group = go[:, model.f2k[f]] == gg[:, model.f2k[f]]
        a, b = vo[same_group, j], vg[same_group, j]
        if model.features[f]["type"] == "nich":
            a, b = a.copy().view(np.float32).astype(np.float64), b.copy().view(np.float32).astype(np.float64)
            assert np.allclose(a, b, rtol=1e-5, atol=1e-5), f
        else:
            assert (a == b).mean() > 0.9995, (f, (a == b).mean())
    assert np.array_equal(qg.query_sample(0, to_sample, 64, seed=7, draw_offset=100), vg[100:164])
    # self-consistency on the device alone: marginals of one feature per discrete type against lb_query_score
    for ftype in ("bb", "dd", "dpd", "gp"):
        j, f = next((j, f) for j, f in enumerate(to_sample) if model.features[f]["type"] == ftype)
        cand = tq.candidates(model, f)
        p = tq.predictive(cuda_abi, model, o, cond, [{f: v} for v in cand])
        counts = np.array([(vg[:, j] == v).sum() for v in cand] + [0.0])
        counts[-1] = n - counts[:-1].sum()
        assert tq.chi_square_p(counts, np.append(p, max(1.0 - p.sum(), 1e-12))) > tq.MIN_P, ftype


# ------------------------------------------------------------------------------------------ the product binary
@pytest.mark.gpu
def test_loom_query_on_the_gpu(tmp_path, infer_on_oracle, oracle_abi):  # noqa: F811
    """loom_query linked against libloomb200.so: Score against the oracle's latents, Sample against the server's own
    scores, Entropy of the empty set, ScoreDerivative leaves the latents as they were"""
    import __graft_entry__
    __graft_entry__.build()
    exe = os.path.join(HOST, "loom_query")
    needed = subprocess.check_output(["readelf", "-d", exe]).decode()
    assert "libloomb200.so" in needed and "oracle" not in needed
    root = tmp_path / "store"
    model, block, rowids = qs.make_store(root, infer_on_oracle)
    qs.mixed_session_case(exe, root, model, block, oracle_abi, tmp_path)


# ------------------------------------------------------------------------------------------ tare / sparsify on the device
@pytest.mark.gpu
def test_differ_on_the_gpu(cuda_abi):
    """lb_differ_* against the numpy restatement of loom::Differ: histograms, tare, CSR diff bit for bit"""
    from test_tare import differ_case
    differ_case(cuda_abi)


@pytest.mark.gpu
def test_loom_tare_and_loom_sparsify_on_the_gpu(tmp_path):
    """the product binaries on the device, in windows of 200 rows"""
    import __graft_entry__
    __graft_entry__.build()
    from test_differ_process import passes_case
    passes_case(os.path.join(HOST, "loom_tare"), os.path.join(HOST, "loom_sparsify"), tmp_path, 200)


# ------------------------------------------------------------------------------------------ committed fixture
@pytest.mark.gpu
def test_cuda_reproduces_query_and_differ_fixture(cuda_abi):
    """tests/golden/query_differ_small.npz through the C ABI on the device"""
    import json
    import test_golden as tg
    from loom_b200 import Model, RowBlock
    z = np.load(os.path.join(tg.GOLDEN, "engine_small.npz"))
    model = Model(json.loads(str(z["features"])), z["f2k"], kind_py=z["kind_py"])
    rows = RowBlock(int(z["n_rows"]), z["cat8"], z["wide"], z["mask"])
    tg.query_differ_case(cuda_abi, (z, model, rows), False)


# ------------------------------------------------------------------------------------------ score error per feature type
@pytest.mark.gpu
def test_score_error_per_feature_type(oracle_abi, cuda_abi):
    """SURVEY.md 8c, parity check 2: device scores against the float64 oracle within 1e-5 * max(1, |score|) for EVERY
    feature type separately (max over all row x group scores), float moments against a float64 recomputation"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HOST), "..", "benchmarks"))
    import score_error
    report = score_error.measure(cuda_abi, oracle_abi, 5000)
    print(report)
    for t, entry in report.items():
        assert entry["max_err"] <= 1e-5, (t, entry)
        assert entry["p999_err"] <= entry["max_err"]
        if "moment_max_rel_err" in entry:
            assert entry["moment_max_rel_err"] <= 1e-9, (t, entry)


# (the 2-GPU NCCL cases live in tests/test_dist.py::test_native_nccl_sharding_two_gpus: the library's own lb_comm_* entry
# points, no torch)


@pytest.mark.gpu
def test_masked_hyper_run_on_one_gpu(oracle_abi, cuda_abi):
    """lb_hyper_run_tasks + lb_set_hypers on one device: two complementary masked runs on two replicas, exchanged by hand,
    end where hyper_run ends (the same property the 2-rank test checks, without a second GPU)"""
    from loom_b200 import Engine, synth
    from test_dpd import GRIDS
    model, rows, truth = synth.generate(1500, [("bb", 3), ("dd16", 2), ("dpd5", 2), ("gp", 3), ("nich", 3)], density=0.6, seed=23)
    grids = dict(GRIDS, bb={"alpha": [0.5, 2.0], "beta": [0.5, 2.0]})

    def engine():
        e = Engine(cuda_abi)
        e.set_model(model)
        e.set_rows(rows)
        for k in range(model.n_kinds):
            uniq, inv = np.unique(truth["assign"][k].astype(np.uint32), return_inverse=True)
            e.set_groups(k, len(uniq))
            e.set_assignments(k, inv.astype(np.uint32))
        e.accumulate_rows()
        e.set_hyper_prior(grids)
        return e
    ref, a, b = engine(), engine(), engine()
    K, F = model.n_kinds, model.n_features
    owner = np.arange(1 + K + F) % 2
    ref.hyper_run(7)
    a.hyper_run(7, task_mask=(owner == 0).astype(np.uint8))
    b.hyper_run(7, task_mask=(owner == 1).astype(np.uint8))
    sa, aa, ka, ta = a.get_hypers()
    sb, ab, kb, tb = b.get_hypers()
    for f in range(F):
        if owner[1 + K + f] == 1:
            for c in range(4):
                sa[f].hp[c] = sb[f].hp[c]
            if sa[f].dim > 0 and sa[f].type in (1, 2):
                aa[sa[f].aux_off:sa[f].aux_off + sa[f].dim] = ab[sa[f].aux_off:sa[f].aux_off + sa[f].dim]
    for k in range(K):
        if owner[1 + k] == 1:
            ka[k] = kb[k]
    a.set_hypers(sa, aa, ka, ta)                    # task 0 (topology) belongs to replica a
    s0, a0, k0, t0 = ref.get_hypers()
    s1, a1, k1, t1 = a.get_hypers()
    for f in range(F):
        assert [s0[f].hp[c] for c in range(4)] == [s1[f].hp[c] for c in range(4)], f
    assert np.array_equal(a0, a1) and np.array_equal(k0, k1) and np.array_equal(t0, t1)
    assert abs(ref.score_data() - a.score_data()) <= 1e-12 * abs(ref.score_data())   # float64 atomics: summation order


# ------------------------------------------------------------------------------------------ the loom.query front end
@pytest.mark.gpu
def test_query_client_on_the_gpu(tmp_path, infer_on_oracle):  # noqa: F811
    """loom_b200.query (loom/query.py's API) over the product binary: the checks of loom/test/test_query_math.py"""
    import __graft_entry__
    __graft_entry__.build()
    from loom_b200 import query
    from test_query_client import typed_row
    root = tmp_path / "store"
    model, block, rowids = qs.make_store(root, infer_on_oracle)
    F = model.n_features
    with query.get_server(root, config={"seed": 2}) as s:          # finds loom_b200/host/loom_query
        assert abs(s.score([None] * F)) < 1e-3
        cond = typed_row(model, qs.some_rows(model, block, [1])[0])
        bb = next(f for f in range(F) if model.features[f]["type"] == "bb")
        cond[bb] = None
        n = 2000
        samples = s.sample([f == bb for f in range(F)], cond, n)
        base = s.score(cond)
        p = [np.exp(s.score(cond[:bb] + [v] + cond[bb + 1:]) - base) for v in (False, True)]
        assert abs(sum(p) - 1) < 1e-4
        ones = sum(bool(x[bb]) for x in samples)
        assert qs.chi_square_p([n - ones, ones], p) > qs.MIN_P
        ent = s.entropy([[bb]], [[]], cond, 500)
        exact = -sum(q * np.log(q) for q in p)
        est = ent[frozenset([bb])]
        assert abs(est.mean - exact) < 4.5 * np.sqrt(est.variance) + 1e-6
        assert len(s.score_derivative(cond, [cond, typed_row(model, qs.some_rows(model, block, [2])[0])])) == 2


# ------------------------------------------------------------------------------------------ fused Score (K1 without scores out)
@pytest.mark.gpu
def test_fused_query_score_matches_the_two_kernel_path(oracle_abi, cuda_abi, monkeypatch):
    """LB_QUERY_FUSED=1: K1 reduces each row's scores to their log-sum-exp in registers (an experiment measured beside
    the default path).  Same numbers as score-then-reduce, same 1e-5 against the oracle; kinds with more than 128
    groups fall back to the two-kernel path."""
    from loom_b200 import Engine, synth, sweep_actions
    for n_rows, types in ((3000, [("bb", 4), ("dd16", 3), ("dd256", 2), ("gp", 3), ("nich", 3)]), (700, [("bb", 3), ("nich", 2)])):
        model, rows, _ = synth.generate(n_rows, types, density=0.6, seed=8)
        g, o = Engine(cuda_abi), Engine(oracle_abi)
        for e in (g, o):
            e.set_model(model)
            e.set_rows(rows)
        g.cat_sweep(sweep_actions(n_rows), seed=2)
        for k in range(model.n_kinds):
            counts, ints, flts = g.get_groups(k)
            keep = counts > 0
            o.set_groups(k, int(keep.sum()), counts[keep], ints[:, keep], flts[:, keep])
            g.set_groups(k, int(keep.sum()), counts[keep], ints[:, keep], flts[:, keep])
        monkeypatch.delenv("LB_QUERY_FUSED", raising=False)
        plain = g.query_score()
        monkeypatch.setenv("LB_QUERY_FUSED", "1")
        fused = g.query_score()
        monkeypatch.delenv("LB_QUERY_FUSED", raising=False)
        want = o.query_score().astype(np.float64)
        assert np.allclose(fused, plain, rtol=2e-6, atol=2e-6)
        assert (np.abs(fused - want) <= 1e-5 * np.maximum(1.0, np.abs(want))).all()
        ragged = g.query_score(5, n_rows - 3)                     # unaligned range, default path
        monkeypatch.setenv("LB_QUERY_FUSED", "1")
        assert np.allclose(g.query_score(5, n_rows - 3), ragged, rtol=2e-6, atol=2e-6)
        monkeypatch.delenv("LB_QUERY_FUSED", raising=False)


# ------------------------------------------------------------------------------------------ generate_rows on the device
@pytest.mark.gpu
def test_generate_rows_matches_oracle_on_the_gpu(oracle_abi, cuda_abi):
    """lb_generate_rows against the oracle: the seating is float64 host code on both sides (identical), the value
    chains use the same Philox counters and float64 samplers (a draw within rounding of a threshold may differ, and a
    (group, feature) chain that differed once stays different), and the device's books balance on their own"""
    import test_generate as tg
    model, g = tg.generated(cuda_abi, n_rows=500, density=0.7, seed=3)
    _, o = tg.generated(oracle_abi, n_rows=500, density=0.7, seed=3)
    for k in range(model.n_kinds):
        assert np.array_equal(g.get_assignments(k), o.get_assignments(k))
        assert np.array_equal(g.get_groups(k)[0], o.get_groups(k)[0])
    rg, ro = g.get_rows(), o.get_rows()
    assert (rg.observed() == ro.observed()).mean() > 0.999
    both = rg.observed() & ro.observed()
    F8 = model.n8()
    for f in range(model.n_features):
        a = (rg.cat8[f] if f < F8 else rg.wide[f - F8])[both[f]]
        b = (ro.cat8[f] if f < F8 else ro.wide[f - F8])[both[f]]
        if model.features[f]["type"] == "nich":
            a, b = a.copy().view(np.float32).astype(np.float64), b.copy().view(np.float32).astype(np.float64)
            assert np.isclose(a, b, rtol=1e-4, atol=1e-4).mean() > 0.98, f
        else:
            assert (a == b).mean() > 0.98, (f, (a == b).mean())
    tg.books_balance(cuda_abi, model, g)


@pytest.mark.gpu
def test_loom_generate_and_loom_mix_on_the_gpu(tmp_path, oracle_abi):
    """the product binaries: a dataset synthesised on the device and mixed, every dumped statistic recomputed by the oracle"""
    import __graft_entry__
    __graft_entry__.build()
    from loom_b200 import RowBlock, pbio, synth
    from test_pbio import GRIDS
    from test_z_loom_infer import TYPES, check_outputs
    model, _, _ = synth.generate(10, TYPES, density=1.0, seed=2)
    pbio.ModelFile.create(model, GRIDS).write(tmp_path / "init.pb.gz")
    c = pbio.config_defaults()
    c.seed, c.generate_row_count, c.generate_density, c.generate_sample_skip = 21, 300, 0.6, 2
    c.kind_iterations, c.kind_empty_kind_count, c.hyper_run = 2, 2, 1
    pbio.config_write(tmp_path / "config.pb.gz", c)
    out = {"model": str(tmp_path / "gen.model.pb.gz"), "groups": str(tmp_path / "gen.groups"), "assign": str(tmp_path / "gen.assign.pbs.gz")}
    proc = subprocess.run([os.path.join(HOST, "loom_generate"), str(tmp_path / "config.pb.gz"), str(tmp_path / "init.pb.gz"),
                           str(tmp_path / "rows.pbs.gz"), out["model"], out["groups"], out["assign"]], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    mf = pbio.ModelFile.read(out["model"])
    rows = pbio.rows_read(tmp_path / "rows.pbs.gz", mf)
    cat8, wide, mask = rows.to_columns(mf)
    block = RowBlock(300, cat8, wide, mask)
    check_outputs(oracle_abi, block, rows.rowids, out)
    mixed = {"model": str(tmp_path / "mix.model.pb.gz"), "groups": str(tmp_path / "mix.groups"), "assign": str(tmp_path / "mix.assign.pbs.gz")}
    proc = subprocess.run([os.path.join(HOST, "loom_mix"), str(tmp_path / "config.pb.gz"), str(tmp_path / "rows.pbs.gz"), out["model"],
                           out["groups"], out["assign"], mixed["model"], mixed["groups"], mixed["assign"]], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    check_outputs(oracle_abi, block, rows.rowids, mixed)
