"""DirichletProcessDiscrete: the hyper-parameter sampler (src/hyper_kernel.cc:90-181) and the
dictionary built at ingest.

The product's sampler is host code inside libloomb200.so (loom_b200/csrc/lb_dpd_hyper.h, called by
lb_hyper_run after it downloads the feature's count table); here it is compiled for the CPU through
tests/dpd_hyper_shim.cc and compared with the oracle's independent C restatement on the same
tables and the same Philox draws.  Statistical checks pin both to the model: the auxiliary table
counts follow the Antoniak law, the resampled betas follow the counts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from loom_b200 import Engine, _abi, synth, sweep_actions
from loom_b200.engine import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory, oracle_abi):
    so = tmp_path_factory.mktemp("shim") / "libdpd_hyper_shim.so"
    oracle_dir = os.path.join(REPO, "oracle")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall",
                           os.path.join(REPO, "tests", "dpd_hyper_shim.cc"), "-o", str(so),
                           "-L" + oracle_dir, "-lloom_oracle", "-Wl,-rpath," + oracle_dir])
    lib = C.CDLL(str(so))
    lib.dpd_hyper_shim.restype = C.c_uint32
    lib.dpd_hyper_shim.argtypes = [_abi._u32p, C.c_int, C.c_int, C.c_uint64, _abi._f32p, _abi._f32p, _abi._f32p, C.c_int,
                                   _abi._f32p, C.c_int, C.c_uint64, C.c_uint64]
    return lib


GRIDS = {"dpd": {"gamma": [0.25, 1.0, 4.0, 16.0], "alpha": [0.5, 2.0, 8.0, 32.0]}}


def dpd_engine(oracle_abi, n_rows=400, dims=(5, 9), seed=0):
    types = [("bb", 2)] + [("dpd%d" % d, 1) for d in dims] + [("nich", 1)]
    model, rows, _ = synth.generate(n_rows, types, density=0.9, seed=seed)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    e.cat_sweep(sweep_actions(n_rows), seed=seed + 1)
    return model, rows, e


def feature_table(e, model, f):
    k = int(model.f2k[f])
    counts, ints, _ = e.get_groups(k)
    off = sum(Model.int_rows(model.features[g]) for g in model.kind_features(k) if g < f)
    dim = len(model.features[f]["betas"])
    return np.ascontiguousarray(ints[off:off + dim + 1]), dim


def run_shim(shim, table, dim, hp, betas, seed, task, grids=GRIDS):
    hp = np.asarray(hp, np.float32).copy()
    betas = np.asarray(betas, np.float32).copy()
    gg = np.asarray(grids["dpd"].get("gamma", []), np.float32)
    ag = np.asarray(grids["dpd"].get("alpha", []), np.float32)
    draws = shim.dpd_hyper_shim(_abi.ptr(table, _abi._u32p), dim, table.shape[1], table.shape[1], _abi.ptr(hp, _abi._f32p),
                                _abi.ptr(betas, _abi._f32p), _abi.ptr(gg, _abi._f32p) if len(gg) else None, len(gg),
                                _abi.ptr(ag, _abi._f32p) if len(ag) else None, len(ag), seed, task)
    return hp, betas, draws


def test_product_sampler_equals_oracle_draw_for_draw(oracle_abi, shim):
    model, rows, e = dpd_engine(oracle_abi)
    dpd = [f for f, ft in enumerate(model.features) if ft["type"] == "dpd"]
    e.set_hyper_prior(GRIDS)
    moved = 0
    for seed in range(6):
        K = e.n_kinds()
        before_specs, before_aux, _, _ = e.get_hypers()
        before = {f: ([before_specs[f].hp[i] for i in range(3)],
                      before_aux[before_specs[f].aux_off:before_specs[f].aux_off + before_specs[f].dim].copy()) for f in dpd}
        tables = {f: feature_table(e, model, f) for f in dpd}
        e.hyper_run(seed=100 + seed)
        specs, aux, _, _ = e.get_hypers()
        for f in dpd:
            table, dim = tables[f]
            hp, betas, draws = run_shim(shim, table, dim, before[f][0], before[f][1], 100 + seed, 1 + K + f)
            got_hp = [specs[f].hp[i] for i in range(3)]
            assert got_hp[0] == hp[0] and got_hp[1] == hp[1]                  # the same grid points
            assert np.allclose(aux[specs[f].aux_off:specs[f].aux_off + dim], betas, rtol=1e-6, atol=1e-9)
            assert abs(got_hp[2] - hp[2]) <= 1e-6
            assert draws > table[:dim].sum() - np.count_nonzero(table[:dim])   # one Bernoulli per customer after the first
            assert abs(betas.sum() + hp[2] - 1.0) < 1e-5 and (betas > 0).all()
            moved += (got_hp[0] != before[f][0][0]) + (got_hp[1] != before[f][0][1])
    assert moved >= 3           # the sampler does move gamma / alpha over a few sweeps


def test_unobserved_value_freezes_the_hypers(oracle_abi, shim):
    """hyper_kernel.cc:124: only infer when every dictionary value has been observed"""
    table = np.zeros((4, 3), np.uint32)            # dim 3, 3 groups; value 2 never observed
    table[0] = [5, 0, 2]
    table[1] = [1, 7, 0]
    table[3] = table[:3].sum(axis=0)
    hp, betas, draws = run_shim(shim, table, 3, [1.0, 2.0, 0.25], [0.25, 0.25, 0.25], 1, 9)
    assert hp.tolist() == [1.0, 2.0, 0.25] and betas.tolist() == [0.25, 0.25, 0.25]
    assert draws == (5 - 1) + (2 - 1) + (7 - 1)    # the auxiliary draws still happen, as in the reference


def test_auxiliary_counts_follow_the_antoniak_law(shim):
    """one cell of n customers, concentration a: E[m] = sum_i a / (a + i).  Read m off the gamma score: with a
    one-point alpha grid and V = 1 the sampler is deterministic given m, so recover the law from many seeds
    through the betas' expectation instead: E[beta_1] = E[m / (m + gamma)] under Dirichlet(m, gamma)."""
    n, a_alpha, beta = 40, 2.0, 0.5
    a = a_alpha * beta
    table = np.asarray([[n], [n]], np.uint32)
    grids = {"dpd": {"gamma": [3.0], "alpha": [a_alpha]}}
    got = []
    for seed in range(3000):
        hp, betas, _ = run_shim(shim, table, 1, [3.0, a_alpha, 1 - beta], [beta], seed, 7, grids)
        got.append(betas[0])
    # exact expectation: sum_m P(m) m / (m + gamma), P(m) from the CRP recursion
    p = np.zeros(n + 1)
    p[1] = 1.0
    for i in range(1, n):
        q = a / (a + i)
        p[1:] = p[1:] * (1 - q) + p[:-1] * q
    want = sum(p[m] * m / (m + 3.0) for m in range(1, n + 1))
    assert abs(np.mean(got) - want) < 4 * np.std(got) / np.sqrt(len(got))


def test_hyper_kernel_leaves_the_posterior_invariant(oracle_abi):
    """a long alternation of cat sweeps and hyper runs keeps betas a distribution and the score finite"""
    model, rows, e = dpd_engine(oracle_abi, n_rows=150, dims=(4,), seed=2)
    e.set_hyper_prior(dict(GRIDS, clustering=[(1.0, 0.1), (4.0, 0.2)]))
    acts = sweep_actions(150, add=True, remove=True)
    for it in range(10):
        e.cat_sweep(acts, seed=50 + it)
        e.hyper_run(seed=80 + it)
        specs, aux, _, _ = e.get_hypers()
        f = [i for i, ft in enumerate(model.features) if ft["type"] == "dpd"][0]
        betas = aux[specs[f].aux_off:specs[f].aux_off + specs[f].dim]
        assert abs(betas.sum() + specs[f].hp[2] - 1) < 1e-5 and (betas > 0).all()
        assert np.isfinite(e.score_data())


# ------------------------------------------------------------------------------------------ dictionaries at ingest
def empty_dpd_model(n_dpd=2, gamma=(0.7, 3.0)):
    feats = [{"type": "bb", "alpha": 0.5, "beta": 0.5}]
    feats += [{"type": "dpd", "gamma": gamma[i], "alpha": 2.0, "beta0": 1.0, "betas": []} for i in range(n_dpd)]
    feats += [{"type": "gp", "alpha": 2.0, "inv_beta": 0.5}]
    return Model(feats, [0] * len(feats))


def write_rows(path, mf, table, rowids=None):
    """table: rows of raw values / None, written through the native writer after a manual CSR build"""
    from loom_b200 import pbio
    ptr, feats, vals = [0], [], []
    for row in table:
        for f, v in enumerate(row):
            if v is not None:
                feats.append(f)
                vals.append(int(v))
        ptr.append(len(feats))
    R = len(table)
    rows = pbio.Rows(np.arange(R) if rowids is None else rowids, np.zeros(R, np.uint8), ptr, feats, vals,
                     np.zeros(R + 1, np.uint64), [])
    return rows


def test_open_dictionaries_follow_the_stream(tmp_path, oracle_abi):
    """loom's init models carry empty DPD dictionaries (loom/generate.py:103-107); values enter in order of
    first appearance, each taking beta0 * Beta(1, gamma) off the stick (Philox stream 4)"""
    import pb_schema
    from loom_b200 import pbio
    from test_pbio import M, fill_value
    model = empty_dpd_model()
    mf = pbio.ModelFile.create(model, None, [[], []])
    assert [len(d) for d in mf.dpd_values()] == [0, 0]
    rng = np.random.default_rng(0)
    raw1 = np.asarray([900, 17, 4000000000, 5, 33])
    raw2 = np.asarray([1, 0])
    msgs, table = [], []
    for r in range(300):
        row = [int(rng.integers(2)), int(raw1[rng.integers(5)]), int(raw2[rng.integers(2)]), int(rng.poisson(2.0))]
        if rng.random() < 0.3:
            row[1] = None
        table.append(row)
        msg = M()["Row"]()
        msg.id = r
        fill_value(msg.diff.pos, model, {f: v for f, v in enumerate(row) if v is not None}, "auto")
        msg.diff.neg.observed.sparsity = 0
        msgs.append(msg)
    path = tmp_path / "rows.pbs.gz"
    pb_schema.dump_stream(path, msgs)
    with pytest.raises(_abi.LoomError, match="dictionary"):
        pbio.rows_read(path, mf)                                   # closed dictionaries: loud
    rows = pbio.rows_read(path, mf, open_dictionaries=True, seed=11, n_threads=4)
    want = [[], []]
    for row in table:
        for j, f in enumerate((1, 2)):
            if row[f] is not None and row[f] not in want[j]:
                want[j].append(row[f])
    got = [d.tolist() for d in mf.dpd_values()]
    assert got == want
    # cells hold dictionary indices
    index = [{v: i for i, v in enumerate(d)} for d in got]
    flat = [(f, (index[f - 1][v] if f in (1, 2) else v)) for row in table for f, v in enumerate(row) if v is not None]
    assert rows.pos_feature.tolist() == [f for f, _ in flat] and rows.pos_value.tolist() == [v for _, v in flat]
    # betas: the stick-breaking recursion on the oracle's Philox
    lo = oracle_abi.lib.lo_uniform
    lo.restype, lo.argtypes = C.c_double, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
    m2 = mf.model()
    for j, f in enumerate((1, 2)):
        gamma, beta0, betas = model.features[f]["gamma"], 1.0, []
        for d in range(len(got[j])):
            u = lo(11, 4, f, d)
            b = max(beta0 * (1.0 - (1.0 - u) ** (1.0 / gamma)), 1e-6)
            beta0 = max(beta0 - b, 1e-6)
            betas.append(b)
        assert np.allclose(m2.features[f]["betas"], betas, rtol=1e-6)
        assert abs(m2.features[f]["beta0"] - beta0) < 1e-6 and abs(sum(betas) + beta0 - 1) < 1e-5
    # a second pass finds nothing new and changes nothing
    before = mf.model().features
    again = pbio.rows_read(path, mf, open_dictionaries=True, seed=99)
    assert [d.tolist() for d in mf.dpd_values()] == want and mf.model().features == before
    assert np.array_equal(again.pos_value, rows.pos_value)
    # another sample adopts the value order with its own betas
    other = pbio.ModelFile.create(empty_dpd_model(gamma=(5.0, 0.2)), None, [[], []])
    other.adopt_dictionaries(mf, seed=12)
    assert [d.tolist() for d in other.dpd_values()] == want
    assert not np.allclose(other.model().features[1]["betas"], m2.features[1]["betas"])
    # the written model is a valid CrossCat message carrying the dictionary
    mf.write(tmp_path / "model.pb.gz")
    g = pb_schema.load_message(tmp_path / "model.pb.gz", M()["CrossCat"])
    assert list(g.kinds[0].product_model.dpd[0].values) == want[0]
    assert len(g.kinds[0].product_model.dpd[1].betas) == 2


def test_unobserved_dpd_feature_gets_one_value(tmp_path):
    from loom_b200 import pbio
    model = empty_dpd_model()
    mf = pbio.ModelFile.create(model, None, [[], []])
    table = [[1, 7, None, 3], [0, 8, None, 1]]
    pbio.rows_write(tmp_path / "rows.pbs", pbio.ModelFile.create(
        Model([dict(f, betas=[0.1] * 9, beta0=0.1) if f["type"] == "dpd" else f for f in model.features],
              [0] * 4)), write_rows(None, None, table))
    pbio.rows_read(tmp_path / "rows.pbs", mf, open_dictionaries=True, seed=1)
    assert [d.tolist() for d in mf.dpd_values()] == [[7, 8], [0]]
