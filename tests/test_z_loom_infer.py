"""`loom_infer` end to end on files: the process boundary loom.runner uses (src/infer.cc:68-81).

The host code under test (loom_b200/host/loom.hpp + loom_infer.cc + the file codec) is the
product's; on this CPU box it is linked against the oracle's copy of the C ABI
(-D'LB_NAME(x)=lo_##x'), on the GPU box against libloomb200.so.  Inputs are written with
google.protobuf (what a loom front end produces), outputs are read back with it and checked
with the invariants loom's own runner tests use (loom/test/test_runner.py:237-250) plus a
recomputation of every sufficient statistic from the rows and the dumped assignments."""
import os
import subprocess

import numpy as np
import pytest

import pb_schema
from conftest import has_gpu
from loom_b200 import Engine, _abi, pbio, synth, tare
from loom_b200.engine import RowBlock
from test_pbio import GRIDS, M, fill_value, io_lib  # noqa: F401  (io_lib: autouse fixture builds the codec)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(REPO, "loom_b200", "host")
TYPES = [("bb", 4), ("dd16", 3), ("dd256", 1), ("gp", 2), ("nich", 3)]


@pytest.fixture(scope="module")
def infer_on_oracle(tmp_path_factory, oracle_abi):
    """loom_infer.cc built against the oracle's lo_ ABI (tests may link oracle/)"""
    exe = tmp_path_factory.mktemp("bin") / "loom_infer_oracle"
    oracle_dir, io_dir = os.path.join(REPO, "oracle"), os.path.join(REPO, "loom_b200", "io")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-DLB_NAME(x)=lo_##x",
                           os.path.join(HOST, "loom_infer.cc"), "-o", str(exe),
                           "-L" + oracle_dir, "-lloom_oracle", "-L" + io_dir, "-lloomb200_io",
                           "-Wl,-rpath," + oracle_dir, "-Wl,-rpath," + io_dir])
    return str(exe)


def write_inputs(d, n_rows=300, seed=0, extra_passes=2.0, kind_iterations=0, hyper=False, tared=False,
                 checkpoint_period_sec=1e9):
    model, block, _ = synth.generate(n_rows, TYPES, density=0.7, seed=seed)
    mf = pbio.ModelFile.create(model, GRIDS if hyper else None)
    mf.write(d / "init.pb.gz")
    rng = np.random.default_rng(seed)
    rowids = rng.permutation(10 * n_rows)[:n_rows].astype(np.uint64)      # arbitrary, shuffled ids
    paths = {"tares": "--none"}
    if tared:
        tare_obs, tare_val = tare.make_tare(model, block)
        diff = tare.sparsify(model, block, tare_obs, tare_val)
        rows = pbio.Rows(rowids, np.ones(n_rows, np.uint8), diff["pos_ptr"], diff["pos_feature"], diff["pos_value"],
                         diff["neg_ptr"], diff["neg_feature"])
        pbio.tares_write(d / "tares.pbs.gz", mf, tare_obs, tare_val)
        paths["tares"] = str(d / "tares.pbs.gz")
    else:
        rows = pbio.Rows.from_block(block, rowids)
    pbio.rows_write(d / "rows.pbs.gz", mf, rows)
    c = pbio.config_defaults()
    c.seed, c.extra_passes, c.kind_iterations, c.kind_empty_kind_count = seed, extra_passes, kind_iterations, 3
    c.hyper_run, c.checkpoint_period_sec = int(hyper), checkpoint_period_sec
    c.small_data_size = 50.0
    pbio.config_write(d / "config.pb.gz", c)
    paths.update(config=str(d / "config.pb.gz"), rows=str(d / "rows.pbs.gz"), model_in=str(d / "init.pb.gz"))
    return model, block, rowids, paths


def run(exe, paths, d, groups_in="--none", assign_in="--none", checkpoint_in="--none", tag="out", checkpoint_out=False,
        expect_failure=False):
    out = {"model": str(d / (tag + ".model.pb.gz")), "groups": str(d / (tag + ".groups")),
           "assign": str(d / (tag + ".assign.pbs.gz")), "log": str(d / (tag + ".log.pbs.gz")),
           "checkpoint": str(d / (tag + ".checkpoint.pb.gz")) if checkpoint_out else "--none"}
    cmd = [exe, paths["config"], paths["rows"], paths["tares"], paths["model_in"], groups_in, assign_in, checkpoint_in,
           out["model"], out["groups"], out["assign"], out["checkpoint"], out["log"]]
    proc = subprocess.run(cmd, capture_output=True, timeout=600)
    if expect_failure:
        assert proc.returncode != 0
        return proc.stderr.decode()
    assert proc.returncode == 0, proc.stderr.decode()
    return out


def check_outputs(oracle_abi, block, rowids, out, n_assigned=None, kinds_fixed_to=None, shuffled=False):
    """structural invariants + every dumped statistic recomputed from rows and assignments"""
    R = block.n_rows
    n_assigned = R if n_assigned is None else n_assigned
    mf = pbio.ModelFile.read(out["model"])
    model = mf.model()
    g = pb_schema.load_message(out["model"], M()["CrossCat"])
    assert g.IsInitialized() and len(g.kinds) == model.n_kinds
    assert sorted(f for k in g.kinds for f in k.featureids) == list(range(model.n_features))
    if kinds_fixed_to is not None:
        assert np.array_equal(model.f2k, kinds_fixed_to)
    a_rowids, a_gids = pbio.assign_read(out["assign"])
    assert len(a_rowids) == n_assigned and a_gids.shape == (n_assigned, model.n_kinds)
    assert len(set(a_rowids.tolist())) == n_assigned and set(a_rowids.tolist()) <= set(rowids.tolist())
    position = {int(r): i for i, r in enumerate(rowids)}
    pos = np.asarray([position[int(r)] for r in a_rowids])
    if n_assigned and not shuffled:
        assert np.array_equal(pos, (pos[0] + np.arange(n_assigned)) % R)      # a window of the cyclic stream, FIFO order
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(block)
    dumped = []
    for k in range(model.n_kinds):
        fids = model.kind_features(k)
        counts, ints, flts = pbio.groups_read(pbio.mixture_path(out["groups"], k), mf, fids)
        assert (counts > 0).all() and counts.sum() == n_assigned
        assert (np.diff(counts.astype(np.int64)) <= 0).all()                  # largest group first (cross_cat.cc:212-236)
        assert a_gids[:, k].max(initial=0) < max(len(counts), 1)
        assert np.array_equal(np.bincount(a_gids[:, k], minlength=len(counts)), counts)
        packed = np.full(R, _abi.UNASSIGNED, np.uint32)
        packed[pos] = a_gids[:, k]
        e.set_groups(k, len(counts))
        e.set_assignments(k, packed)
        dumped.append((counts, ints, flts))
    e.accumulate_rows()
    for k, (counts, ints, flts) in enumerate(dumped):
        c2, i2, f2 = e.get_groups(k)
        n = len(counts)
        assert np.array_equal(c2[:n], counts) and np.array_equal(i2[:, :n], ints)       # integers bit-exact
        assert np.allclose(f2[:, :n], flts, rtol=2e-5, atol=2e-5)                       # float32 on the wire
    return model, a_rowids, a_gids


def test_multi_pass_cat_only(tmp_path, infer_on_oracle, oracle_abi):
    model, block, rowids, paths = write_inputs(tmp_path)
    out = run(infer_on_oracle, paths, tmp_path)
    check_outputs(oracle_abi, block, rowids, out, kinds_fixed_to=model.f2k)
    logs = pb_schema.load_stream(out["log"], M()["LogMessage"])
    assert len(logs) >= 3 and all(m.IsInitialized() for m in logs)
    assert [m.args.iter for m in logs] == list(range(len(logs)))
    assert logs[-1].args.scores.assigned_object_count == block.n_rows
    assert list(logs[-1].args.summary.feature_counts) == [len(model.kind_features(k)) for k in range(model.n_kinds)]
    assert np.isfinite(logs[-1].args.scores.score) and logs[-1].args.scores.score < 0
    # same seed, same files -> same result; another seed -> another sample
    again = run(infer_on_oracle, paths, tmp_path, tag="again")
    assert np.array_equal(pbio.assign_read(out["assign"])[1], pbio.assign_read(again["assign"])[1])


def test_single_pass(tmp_path, infer_on_oracle, oracle_abi):
    model, block, rowids, paths = write_inputs(tmp_path, extra_passes=0.0)
    out = run(infer_on_oracle, paths, tmp_path)
    # infer_single_pass streams assignments in row order with the ids given at assignment time
    a_rowids, a_gids = pbio.assign_read(out["assign"])
    assert np.array_equal(a_rowids, rowids) and a_gids.shape == (block.n_rows, model.n_kinds)
    mf = pbio.ModelFile.read(out["model"])
    for k in range(model.n_kinds):
        counts, _, _ = pbio.groups_read(pbio.mixture_path(out["groups"], k), mf, model.kind_features(k))
        assert counts.sum() == block.n_rows
        assert sorted(np.bincount(a_gids[:, k]).tolist(), reverse=True)[:len(counts)] == counts.tolist()


def test_kind_and_hyper_kernels_and_tares(tmp_path, infer_on_oracle, oracle_abi):
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=200, seed=3, kind_iterations=4, hyper=True, tared=True)
    out = run(infer_on_oracle, paths, tmp_path)
    new_model, _, _ = check_outputs(oracle_abi, block, rowids, out)
    g = pb_schema.load_message(out["model"], M()["CrossCat"])
    assert g.hyper_prior == pb_schema.load_message(paths["model_in"], M()["CrossCat"]).hyper_prior   # grids pass through
    # hypers moved onto grid points
    assert all(any(np.isclose(k.product_model.clustering.alpha, a) for a, _ in GRIDS["clustering"]) for k in g.kinds)
    for f in new_model.features:
        if f["type"] == "bb":
            assert any(np.isclose(f["alpha"], a) for a in GRIDS["bb"]["alpha"])
    logs = pb_schema.load_stream(out["log"], M()["LogMessage"])
    assert logs[-1].args.kernel_status.kind.total_count >= 1


def test_checkpoint_resume_loop(tmp_path, infer_on_oracle, oracle_abi):
    """checkpoint_period_sec = 0: every call stops at its first batch boundary; feeding the outputs back
    in (groups, assignments, checkpoint) until `finished` is loom's resume protocol (src/loom.cc:178-214)"""
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=120, seed=5, checkpoint_period_sec=0.0)
    groups_in = assign_in = checkpoint_in = "--none"
    model_in = paths["model_in"]
    seen_iters, assigned = [], []
    for step in range(200):
        p = dict(paths, model_in=model_in)
        out = run(infer_on_oracle, p, tmp_path, groups_in, assign_in, checkpoint_in, tag="step%d" % step,
                  checkpoint_out=True)
        ck = pbio.checkpoint_read(out["checkpoint"])
        g = pb_schema.load_message(out["checkpoint"], M()["Checkpoint"])
        assert g.IsInitialized() and g.row_count == 120 and g.tardis_iter == ck.tardis_iter
        seen_iters.append(ck.tardis_iter)
        n = len(pbio.assign_read(out["assign"])[0])
        assigned.append(n)
        assert (ck.unassigned_pos - ck.assigned_pos) % 120 == n % 120
        if ck.finished:
            break
        check_outputs(oracle_abi, block, rowids, out, n_assigned=n, kinds_fixed_to=model.f2k)
        groups_in, assign_in, checkpoint_in, model_in = out["groups"], out["assign"], out["checkpoint"], out["model"]
    assert ck.finished and assigned[-1] == 120 and len(seen_iters) > 3
    assert all(b > a for a, b in zip(seen_iters, seen_iters[1:]))
    check_outputs(oracle_abi, block, rowids, out, kinds_fixed_to=model.f2k)
    logs = pb_schema.load_stream(out["log"], M()["LogMessage"])      # each resumed call appends to its own log
    assert len(logs) >= 1


def test_errors_abort_with_looms_message(tmp_path, infer_on_oracle):
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=50)
    err = run(infer_on_oracle, dict(paths, rows=str(tmp_path / "missing.pbs.gz")), tmp_path, expect_failure=True)
    assert err.startswith("ERROR failed to open input file") and "loom.hpp" in err      # ERROR msg \n\t file : line \n\t function
    err = run(infer_on_oracle, dict(paths, model_in=paths["rows"]), tmp_path, expect_failure=True)
    assert err.startswith("ERROR")
    # assignments that are not a window of the stream
    out = run(infer_on_oracle, paths, tmp_path)
    a_rowids, a_gids = pbio.assign_read(out["assign"])
    pbio.assign_write(tmp_path / "scrambled.pbs.gz", a_rowids[::-1][:20], a_gids[::-1][:20])
    err = run(infer_on_oracle, paths, tmp_path, groups_in=out["groups"], assign_in=str(tmp_path / "scrambled.pbs.gz"),
              expect_failure=True)
    assert "contiguous window" in err or "sample_size" in err
    proc = subprocess.run([infer_on_oracle, "too", "few"], capture_output=True)
    assert proc.returncode == 1 and b"Usage: loom_infer CONFIG_IN ROWS_IN TARES_IN MODEL_IN" in proc.stderr


def test_product_binary_links_against_the_cuda_library():
    """`make -C loom_b200/host` links loom_infer with libloomb200.so (no oracle anywhere near it)"""
    import __graft_entry__
    __graft_entry__.build()
    exe = os.path.join(HOST, "loom_infer")
    assert os.path.exists(exe)
    needed = subprocess.check_output(["readelf", "-d", exe]).decode()
    assert "libloomb200.so" in needed and "libloomb200_io.so" in needed and "oracle" not in needed
    if not has_gpu():
        # without a device the engine refuses to start: there is no CPU fallback behind the binary
        proc = subprocess.run([exe] + ["x"] * 12, capture_output=True)
        assert proc.returncode != 0


@pytest.mark.gpu
def test_loom_infer_on_the_gpu(tmp_path, oracle_abi):
    """the product binary on the device: files in, files out, statistics recomputed by the oracle"""
    import __graft_entry__
    __graft_entry__.build()
    exe = os.path.join(HOST, "loom_infer")
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=400, seed=7, kind_iterations=4, hyper=True, tared=True)
    out = run(exe, paths, tmp_path)
    check_outputs(oracle_abi, block, rowids, out)
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=300, seed=8)
    out = run(exe, paths, tmp_path, tag="cat")
    check_outputs(oracle_abi, block, rowids, out, kinds_fixed_to=model.f2k)


# ------------------------------------------------------------------------------------------ a row order per chain
def shuffled_streams(n_rows, n_chains, seed):
    """[n_chains][3R] actions: ADD every row, then REMOVE/ADD every row, each chain in its own row order
    (loom.runner.shuffle with the sample's seed, loom/tasks.py:228-233); same ADD/REMOVE pattern for all"""
    rng = np.random.default_rng(seed)
    out = np.empty((n_chains, 3 * n_rows), np.uint32)
    for c in range(n_chains):
        perm = rng.permutation(n_rows).astype(np.uint32)
        out[c, :n_rows] = perm << 1
        out[c, n_rows::2] = (perm << 1) | 1
        out[c, n_rows + 1::2] = perm << 1
    return out


def streams_case(abi, n_chains=11, n_rows=500):
    from loom_b200 import cat_sweep_streams
    model, block, _ = synth.generate(n_rows, TYPES, density=0.6, seed=4)
    actions = shuffled_streams(n_rows, n_chains, seed=9)
    seeds = np.arange(100, 100 + n_chains)
    owner = Engine(abi)
    owner.set_model(model)
    owner.set_rows(block)
    chains = [owner]
    for _ in range(n_chains - 1):
        e = Engine(abi)
        e.set_model(model)
        e.share_rows(owner)
        chains.append(e)
    cat_sweep_streams(chains, actions, seeds)
    return model, block, actions, seeds, chains


def test_streams_equal_single_chain_sweeps_oracle(oracle_abi):
    model, block, actions, seeds, chains = streams_case(oracle_abi, n_chains=3, n_rows=200)
    for c, chain in enumerate(chains):
        single = Engine(oracle_abi)
        single.set_model(model)
        single.set_rows(block)
        single.cat_sweep(actions[c], seed=int(seeds[c]))
        for k in range(model.n_kinds):
            assert np.array_equal(single.get_assignments(k), chain.get_assignments(k))
            assert all(np.array_equal(x, y) for x, y in zip(single.get_groups(k), chain.get_groups(k)))
    assert not np.array_equal(chains[0].get_assignments(0), chains[1].get_assignments(0))


@pytest.mark.gpu
@pytest.mark.parametrize("warps", [1, 8])
def test_streams_equal_single_chain_sweeps_gpu(oracle_abi, cuda_abi, monkeypatch, warps):
    """lb_cat_sweep_streams: one launch, every chain in its own row order over the shared block; each chain
    ends bit for bit where its own single-chain sweep ends, and one of them replays on the oracle.  With one
    warp per (chain, kind) eight chains share a block and re-align at every action while reading different rows."""
    import test_cuda_cat as t
    monkeypatch.setenv("LB_SWEEP_WARPS", str(warps))
    model, block, actions, seeds, chains = streams_case(cuda_abi)
    for c, chain in enumerate(chains):
        single = Engine(cuda_abi)
        single.set_model(model)
        single.set_rows(block)
        picks, scores = single.cat_sweep(actions[c], seed=int(seeds[c]), debug=True, debug_stride=64)
        if c == 0:
            o = Engine(oracle_abi)
            o.set_model(model)
            o.set_rows(block)
            bad, err = t.replay(oracle_abi, o, actions[c], int(seeds[c]), picks, scores, 64)
            assert bad == 0 and err <= 1e-5
            t.assert_same_state(o, single, model)
        for k in range(model.n_kinds):
            assert np.array_equal(single.get_assignments(k), chain.get_assignments(k))
            assert all(np.array_equal(x, y) for x, y in zip(single.get_groups(k), chain.get_groups(k)))
    assert not np.array_equal(chains[0].get_assignments(0), chains[1].get_assignments(0))


# ------------------------------------------------------------------------------------------ loom.tasks.infer: many samples
def tasks_case(abi, tmp_path, oracle_abi, n_samples=3):
    from loom_b200 import hyperprior, tasks
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=150, seed=6, tared=True)
    feature_sizes = [{"type": "dd", "dim": len(f["alphas"])} if f["type"] == "dd" else {"type": f["type"]}
                     for f in model.features]
    inits = []
    for s in range(n_samples):                      # a fresh init per sample (loom/tasks.py:222-226)
        init = tasks.generate_init(feature_sizes, hyperprior.defaults(), seed=s)
        path = tmp_path / ("init.%d.pb.gz" % s)
        pbio.ModelFile.create(init, GRIDS).write(path)
        inits.append(str(path))
    assert len({tuple(pbio.ModelFile.read(p).model().f2k.tolist()) for p in inits}) > 1
    outs = [{"model": str(tmp_path / ("s%d.model.pb.gz" % s)), "groups": str(tmp_path / ("s%d.groups" % s)),
             "assign": str(tmp_path / ("s%d.assign.pbs.gz" % s))} for s in range(n_samples)]
    config = {"schedule": {"extra_passes": 2.0, "small_data_size": 50.0}, "kernels": {"kind": {"iterations": 3, "empty_kind_count": 2}}}
    tasks.infer(paths["rows"], inits, outs, config, tares_in=paths["tares"], abi=abi)
    seen = []
    for out in outs:
        m, a_rowids, a_gids = check_outputs(oracle_abi, block, rowids, out, shuffled=True)
        seen.append((tuple(a_rowids[:10].tolist()), m.n_kinds))
    assert len({s[0] for s in seen}) == n_samples       # every sample walked the rows in its own order
    return outs


def test_tasks_infer_many_samples_oracle(tmp_path, oracle_abi):
    tasks_case(oracle_abi, tmp_path, oracle_abi)


@pytest.mark.gpu
def test_tasks_infer_many_samples_gpu(tmp_path, oracle_abi, cuda_abi):
    tasks_case(cuda_abi, tmp_path, oracle_abi, n_samples=9)


# ------------------------------------------------------------------------------------------ DPD: open dictionaries
DPD_TYPES = [("bb", 3), ("dd16", 1), ("dpd6", 2), ("gp", 1), ("nich", 2)]


def raw_of(feature, index):
    """raw value a front end would have given the index-th category of DPD feature `feature`"""
    return 1000 * feature + 7 * int(index) + 3


def dpd_case(exe, tmp_path, oracle_abi, n_rows=250, seed=9):
    from loom_b200.engine import Model
    model, block, _ = synth.generate(n_rows, DPD_TYPES, density=0.8, seed=seed)
    dpd_feats = [f for f, ft in enumerate(model.features) if ft["type"] == "dpd"]
    # what loom hands over: an init model whose DPD dictionaries are EMPTY, rows holding raw values
    init = Model([dict(ft, betas=[], beta0=1.0) if ft["type"] == "dpd" else ft for ft in model.features], model.f2k)
    grids = dict(GRIDS, dpd={"gamma": [0.5, 2.0], "alpha": [0.5, 2.0, 8.0]})
    pbio.ModelFile.create(init, grids, [[] for _ in dpd_feats]).write(tmp_path / "init.pb.gz")
    writer_model = pbio.ModelFile.create(model, None, [[raw_of(f, i) for i in range(6)] for f in dpd_feats])
    rowids = np.arange(n_rows, dtype=np.uint64) + 50
    pbio.rows_write(tmp_path / "rows.pbs.gz", writer_model, pbio.Rows.from_block(block, rowids))
    c = pbio.config_defaults()
    c.seed, c.extra_passes, c.kind_iterations, c.kind_empty_kind_count, c.small_data_size = seed, 2.0, 2, 2, 50.0
    pbio.config_write(tmp_path / "config.pb.gz", c)
    paths = {"config": str(tmp_path / "config.pb.gz"), "rows": str(tmp_path / "rows.pbs.gz"), "tares": "--none",
             "model_in": str(tmp_path / "init.pb.gz")}
    out = run(exe, paths, tmp_path)
    # the output model carries the dictionary in order of first appearance; put the block into that numbering
    mf = pbio.ModelFile.read(out["model"])
    dicts = [d.tolist() for d in mf.dpd_values()]
    F8 = model.n8()
    remapped = RowBlock(block.n_rows, block.cat8, block.wide.copy(), block.mask)
    obs = block.observed()
    for j, f in enumerate(dpd_feats):
        seen = sorted({int(v) for v in block.wide[f - F8][obs[f]]})
        assert sorted(dicts[j]) == [raw_of(f, i) for i in seen]          # exactly the observed values
        first = []
        for v in block.wide[f - F8][obs[f]]:
            if raw_of(f, v) not in first:
                first.append(raw_of(f, v))
        assert dicts[j] == first                                          # in stream order
        lut = np.zeros(6, np.uint32)
        for i in seen:
            lut[i] = dicts[j].index(raw_of(f, i))
        remapped.wide[f - F8] = np.where(obs[f], lut[block.wide[f - F8]], 0)
    new_model, _, _ = check_outputs(oracle_abi, remapped, rowids, out)
    for f in dpd_feats:
        ft = new_model.features[f]
        assert abs(sum(ft["betas"]) + ft["beta0"] - 1.0) < 1e-4 and min(ft["betas"]) > 0
        assert ft["gamma"] in (0.5, 2.0) and ft["alpha"] in (0.5, 2.0, 8.0)   # the DPD hyper kernel ran


def test_dpd_features_with_empty_init_dictionaries(tmp_path, infer_on_oracle, oracle_abi):
    dpd_case(infer_on_oracle, tmp_path, oracle_abi)


@pytest.mark.gpu
def test_dpd_features_with_empty_init_dictionaries_gpu(tmp_path, oracle_abi):
    import __graft_entry__
    __graft_entry__.build()
    dpd_case(os.path.join(HOST, "loom_infer"), tmp_path, oracle_abi)


# ------------------------------------------------------------------------------------------ loom_posterior_enum
def build_on_oracle(tmp_path_factory, source, name):
    exe = tmp_path_factory.mktemp("bin") / name
    oracle_dir, io_dir = os.path.join(REPO, "oracle"), os.path.join(REPO, "loom_b200", "io")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-DLB_NAME(x)=lo_##x",
                           os.path.join(HOST, source), "-o", str(exe), "-L" + oracle_dir, "-lloom_oracle",
                           "-L" + io_dir, "-lloomb200_io", "-Wl,-rpath," + oracle_dir, "-Wl,-rpath," + io_dir])
    return str(exe)


@pytest.fixture(scope="module")
def posterior_enum_on_oracle(tmp_path_factory, oracle_abi):
    return build_on_oracle(tmp_path_factory, "loom_posterior_enum.cc", "loom_posterior_enum_oracle")


def posterior_enum_case(exe, d, R, C, feature_type, density, infer_kinds, seed, ingest=None):
    """loom/test/test_posterior_enum.py through the process: write the files, run the binary, parse the
    samples stream (parse_sample, :738-745) and return the goodness of fit of the visited latents"""
    import pe_util
    rng = np.random.default_rng(1000 + seed)
    model = pe_util.make_model(C, feature_type)
    table = pe_util.generate_rows(R, C, feature_type, density, rng)
    block = RowBlock.from_table(model, table)
    mf = pbio.ModelFile.create(model)
    mf.write(d / "model.pb.gz")
    pbio.rows_write(d / "rows.pbs.gz", mf, pbio.Rows.from_block(block))
    size = pe_util.LATENT_SIZES[R][C] if infer_kinds else pe_util.LATENT_SIZES[R][1]
    c = pbio.config_defaults()
    c.seed, c.posterior_enum_sample_count, c.posterior_enum_sample_skip = seed, 10 * size, 10
    c.kind_iterations, c.kind_empty_kind_count, c.hyper_run = (32 if infer_kinds else 0), 32, 0
    pbio.config_write(d / "config.pb.gz", c)
    out = d / "samples.pbs.gz"
    rows_in, tares_in = str(d / "rows.pbs.gz"), "--none"
    if ingest:      # the reference's test runs through tare + sparsify by default (loom/test/test_posterior_enum.py:675-710)
        tare_exe, sparsify_exe = ingest
        schema_row = M()["ProductValue"]()
        schema_row.observed.sparsity = 2
        for f in model.features:
            schema_row.observed.dense.append(True)
            (schema_row.booleans if f["type"] == "bb" else schema_row.reals if f["type"] == "nich" else schema_row.counts).append(0)
        pb_schema.dump_message(d / "schema.pb.gz", schema_row)
        for cmd in ([tare_exe, d / "schema.pb.gz", d / "rows.pbs.gz", d / "tares.pbs.gz"],
                    [sparsify_exe, d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "diffs.pbs.gz"]):
            p = subprocess.run([str(c) for c in cmd], capture_output=True, timeout=300)
            assert p.returncode == 0, p.stderr.decode()
        rows_in, tares_in = str(d / "diffs.pbs.gz"), str(d / "tares.pbs.gz")
    proc = subprocess.run([exe, str(d / "config.pb.gz"), rows_in, tares_in, str(d / "model.pb.gz"),
                           "--none", "--none", str(out)], capture_output=True, timeout=900)
    assert proc.returncode == 0, proc.stderr.decode()
    samples = []
    for msg in pb_schema.load_stream(out, M()["PosteriorEnum.Sample"]):
        latent = frozenset((frozenset(k.featureids), frozenset(frozenset(g.rowids) for g in k.groups)) for k in msg.kinds)
        samples.append((latent, msg.score))
    assert len(samples) == 10 * size
    return pe_util.goodness_of_fit(samples, R, C, infer_kinds)


@pytest.mark.parametrize("feature_type", ["bb", "nich"])
def test_posterior_enum_process_cats(tmp_path, posterior_enum_on_oracle, feature_type):
    import pe_util
    for i, (R, C) in enumerate([(3, 2), (4, 1)]):
        d = tmp_path / ("c%d" % i)
        d.mkdir()
        gof, info = posterior_enum_case(posterior_enum_on_oracle, d, R, C, feature_type, 0.5, False, seed=i)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)


def test_posterior_enum_process_kinds(tmp_path, posterior_enum_on_oracle):
    import pe_util
    for i, (R, C) in enumerate([(2, 2), (3, 2)]):
        d = tmp_path / ("k%d" % i)
        d.mkdir()
        gof, info = posterior_enum_case(posterior_enum_on_oracle, d, R, C, "dd", 1.0, True, seed=10 + i)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)


@pytest.mark.gpu
def test_posterior_enum_process_gpu(tmp_path):
    """loom's chi-square test through the product binary on the device"""
    import __graft_entry__
    import pe_util
    __graft_entry__.build()
    exe = os.path.join(HOST, "loom_posterior_enum")
    for i, (R, C, kinds, ftype) in enumerate([(3, 2, False, "gp"), (3, 2, True, "bb")]):
        d = tmp_path / ("g%d" % i)
        d.mkdir()
        gof, info = posterior_enum_case(exe, d, R, C, ftype, 0.5, kinds, seed=20 + i)
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)


# ------------------------------------------------------------------------------------------ DPD hyper kernel on the device
@pytest.mark.gpu
def test_hyper_run_with_dpd_matches_oracle_on_the_gpu(oracle_abi, cuda_abi):
    from test_dpd import GRIDS, dpd_engine
    model, rows, o = dpd_engine(oracle_abi)
    g = Engine(cuda_abi)
    g.set_model(model)
    g.set_rows(rows)
    for k in range(model.n_kinds):              # the oracle's state, so both start from the same tables
        counts, ints, flts = o.get_groups(k)
        n = int(np.count_nonzero(counts))
        g.set_groups(k, n, counts[:n], ints[:, :n], flts[:, :n])
        g.set_assignments(k, o.get_assignments(k))
    grids = dict(GRIDS, bb={"alpha": [0.5, 2.0], "beta": [0.5, 2.0]})
    o.set_hyper_prior(grids)
    g.set_hyper_prior(grids)
    for seed in (1, 2, 3):
        o.hyper_run(seed=seed)
        g.hyper_run(seed=seed)
        so, ao, _, _ = o.get_hypers()
        sg, ag, _, _ = g.get_hypers()
        for f in range(model.n_features):
            assert [so[f].hp[i] for i in range(2)] == [sg[f].hp[i] for i in range(2)], f
        assert np.allclose(ao, ag, rtol=1e-5, atol=1e-8)
        assert g.score_data() == pytest.approx(o.score_data(), rel=1e-5)

