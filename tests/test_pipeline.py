"""The processes chained the way loom.tasks chains them (ingest -> infer -> query), all built against the oracle's lo_ ABI:
loom_tare -> loom_sparsify -> loom_shuffle -> loom_infer (x3 seeds, init models with EMPTY DPD dictionaries, the tares
file) -> a store -> loom_query.  The data has a DPD column whose dominant raw value ends up in the tare row: that value is in no row's
pos, so it has to enter the dictionary from the tares file (lbio_tares_read_open)."""
import shutil

import numpy as np
import pytest

import pb_schema
from loom_b200 import RowBlock, pack_mask, pbio, synth
from loom_b200.engine import Model
from test_pbio import GRIDS, M, fill_value, io_lib  # noqa: F401
from test_differ_process import run as run_cmd, sparsify_on_oracle, tare_on_oracle  # noqa: F401
from test_query_server import (N_SAMPLES, cells_of, fill_diff, new_request, query_on_oracle, serve,  # noqa: F401
                               some_rows, value_cells)
from test_z_loom_infer import (DPD_TYPES, infer_on_oracle, posterior_enum_case, posterior_enum_on_oracle, raw_of,  # noqa: F401
                               run as run_infer)


def raw_dataset(n_rows=300, seed=5):
    """a block in RAW values (what a front end writes) with sparse booleans and a dominant DPD value below 16"""
    rng = np.random.default_rng(seed)
    model, block, _ = synth.generate(n_rows, DPD_TYPES, density=0.85, seed=seed)
    F8 = model.n8()
    dpd = [f for f, ft in enumerate(model.features) if ft["type"] == "dpd"]
    raws = {dpd[0]: [9, 200, 3, 77, 1000, 14], dpd[1]: [raw_of(dpd[1], i) for i in range(6)]}
    block.cat8[:3] = rng.random((3, n_rows)) < 0.06
    block.wide[dpd[0] - F8] = np.where(rng.random(n_rows) < 0.8, 0, block.wide[dpd[0] - F8])     # index 0 <-> raw 9
    obs = block.observed()
    obs[:3] = True
    obs[dpd[0]] |= rng.random(n_rows) < 0.95
    block = RowBlock(n_rows, block.cat8, block.wide, pack_mask(obs))
    return model, block, dpd, raws


def test_ingest_infer_query(tmp_path, tare_on_oracle, sparsify_on_oracle, infer_on_oracle, query_on_oracle, oracle_abi):  # noqa: F811
    d = tmp_path
    model, block, dpd, raws = raw_dataset()
    writer = pbio.ModelFile.create(model, None, [raws[f] for f in dpd])
    rowids = np.arange(block.n_rows, dtype=np.uint64) + 7
    pbio.rows_write(d / "rows.pbs.gz", writer, pbio.Rows.from_block(block, rowids))
    schema_row = M()["ProductValue"]()
    schema_row.observed.sparsity = 2
    for f in model.features:
        schema_row.observed.dense.append(True)
        (schema_row.booleans if f["type"] == "bb" else schema_row.reals if f["type"] == "nich" else schema_row.counts).append(0)
    pb_schema.dump_message(d / "schema.pb.gz", schema_row)
    # ingest
    assert run_cmd([tare_on_oracle, d / "schema.pb.gz", d / "rows.pbs.gz", d / "tares.pbs.gz"]) == (0, "")
    tare = value_cells(model, pb_schema.load_stream(d / "tares.pbs.gz", M()["ProductValue"])[0])
    assert tare[dpd[0]] == 9 and all(tare[f] == 0 for f in range(3))          # RAW value of the DPD column
    assert run_cmd([sparsify_on_oracle, d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "diffs.pbs.gz"]) == (0, "")
    # infer three samples from init models with empty dictionaries
    init = Model([dict(ft, betas=[], beta0=1.0) if ft["type"] == "dpd" else ft for ft in model.features], model.f2k)
    pbio.ModelFile.create(init, GRIDS, [[] for _ in dpd]).write(d / "init.pb.gz")
    root = d / "store"
    (root / "ingest").mkdir(parents=True)
    shutil.copy(d / "diffs.pbs.gz", root / "ingest" / "diffs.pbs.gz")
    shutil.copy(d / "tares.pbs.gz", root / "ingest" / "tares.pbs.gz")
    for s in range(N_SAMPLES):
        c = pbio.config_defaults()
        c.seed, c.extra_passes, c.kind_iterations, c.hyper_run, c.small_data_size = 40, 1.5, 0, 1, 50.0
        pbio.config_write(d / "config.pb.gz", c)
        # loom.tasks.infer_one shuffles the diffs with the sample's seed first (loom/tasks.py:228-233)
        import os
        shuffle_exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "loom_b200", "host", "loom_shuffle")
        assert run_cmd([shuffle_exe, d / "diffs.pbs.gz", d / ("shuffled.%d.pbs.gz" % s), s]) == (0, "")
        paths = {"config": str(d / "config.pb.gz"), "rows": str(d / ("shuffled.%d.pbs.gz" % s)), "tares": str(d / "tares.pbs.gz"),
                 "model_in": str(d / "init.pb.gz")}
        out = run_infer(infer_on_oracle, paths, d, tag="s%d" % s)
        sd = root / "samples" / ("sample.%d" % s)
        sd.mkdir(parents=True)
        shutil.copy(d / "config.pb.gz", sd / "config.pb.gz")
        shutil.copy(out["model"], sd / "model.pb.gz")
        shutil.copytree(out["groups"], sd / "groups")
        a_rowids, a_gids = pbio.assign_read(out["assign"])
        assert sorted(a_rowids.tolist()) == rowids.tolist()
        mf = pbio.ModelFile.read(out["model"])
        dicts = [x.tolist() for x in mf.dpd_values()]
        assert 9 in dicts[0]                                                   # came in through the tares file
        obs = block.observed()
        F8 = model.n8()
        for j, f in enumerate(dpd):
            assert sorted(dicts[j]) == sorted({raws[f][int(v)] for v in block.wide[f - F8][obs[f]]})
        # every sufficient statistic of the dump adds up to the data: counts per DPD value over all groups
        for k in range(mf.model().n_kinds):
            fids = mf.model().kind_features(k)
            counts, ints, flts = pbio.groups_read(pbio.mixture_path(out["groups"], k), mf, fids)
            assert counts.sum() == block.n_rows
            row = 0
            for f in fids:
                ft = mf.model().features[f]
                if ft["type"] == "dpd":
                    j = dpd.index(f)
                    for i, raw in enumerate(dicts[j]):
                        want = int(((block.wide[f - F8] == raws[f].index(raw)) & obs[f]).sum())
                        assert int(ints[row + i].sum()) == want, (f, raw)
                row += {"bb": 2, "gp": 2, "nich": 1}.get(ft["type"]) or (len(ft["alphas"] if ft["type"] == "dd" else ft["betas"]) + 1)
    qc = pbio.config_defaults()
    qc.seed = 3
    pbio.config_write(root / "query_config.pb.gz", qc)
    # query: a row given densely (raw values) and given as its diff against the tare score the same
    diffs = pb_schema.load_stream(d / "diffs.pbs.gz", M()["Row"])
    dense = pb_schema.load_stream(d / "rows.pbs.gz", M()["Row"])
    requests = []
    for i in range(5):
        r = new_request(2 * i)
        r.score.data.CopyFrom(dense[i].diff)
        requests.append(r)
        r = new_request(2 * i + 1)
        r.score.data.CopyFrom(diffs[i].diff)
        requests.append(r)
    s = new_request(100)                                                       # sample the tared DPD column given nothing
    s.sample.data.pos.observed.sparsity = 0
    s.sample.data.neg.observed.sparsity = 0
    s.sample.to_sample.sparsity = 1
    s.sample.to_sample.sparse.append(dpd[0])
    s.sample.sample_count = 400
    requests.append(s)
    responses = serve(query_on_oracle, root, requests, d)
    assert not any(r.error for r in responses)
    scores = [r.score.score for r in responses[:10]]
    assert all(np.isfinite(scores))
    for i in range(5):
        assert scores[2 * i] == pytest.approx(scores[2 * i + 1], abs=1e-5)
    # ... and equals the log-mean over the samples computed here, every sample with its OWN dictionary order (the samples
    # saw the rows in different orders, so the orders differ and the server has to line them up)
    from loom_b200 import Engine
    orders, per_sample = set(), []
    F8 = model.n8()
    obs = block.observed()
    for s_ in range(N_SAMPLES):
        sd = root / "samples" / ("sample.%d" % s_)
        mf = pbio.ModelFile.read(sd / "model.pb.gz")
        sample_model = mf.model()
        dicts = [x.tolist() for x in mf.dpd_values()]
        orders.add(tuple(map(tuple, dicts)))
        table = []
        for r in range(5):
            row = [None] * model.n_features
            for f in range(model.n_features):
                if not obs[f, r]:
                    continue
                t = model.features[f]["type"]
                if f < F8:
                    row[f] = int(block.cat8[f, r])
                elif t == "nich":
                    row[f] = float(block.wide[f - F8, r:r + 1].view(np.float32)[0])
                elif t == "dpd":
                    row[f] = dicts[dpd.index(f)].index(raws[f][int(block.wide[f - F8, r])])
                else:
                    row[f] = int(block.wide[f - F8, r])
            table.append(row)
        e = Engine(oracle_abi)
        e.set_model(sample_model)
        e.set_rows(RowBlock.from_table(sample_model, table))
        for k in range(sample_model.n_kinds):
            counts, ints, flts = pbio.groups_read(pbio.mixture_path(sd / "groups", k), mf, sample_model.kind_features(k))
            e.set_groups(k, len(counts), counts, ints, flts)
        per_sample.append(e.query_score().astype(np.float64))
    assert len(orders) > 1                                                     # the dictionaries really are ordered differently
    per_sample = np.stack(per_sample)
    m = per_sample.max(axis=0)
    want = m + np.log(np.exp(per_sample - m).sum(axis=0)) - np.log(N_SAMPLES)
    assert np.allclose([scores[2 * i] for i in range(5)], want, rtol=1e-5, atol=1e-5)
    drawn = [value_cells(model, x.pos)[dpd[0]] for x in responses[10].sample.samples]
    assert set(drawn) <= set(raws[dpd[0]]) | {0xFFFFFFFF}                      # raw values (or distributions' OTHER)
    assert np.mean([v == 9 for v in drawn]) > 0.5                              # the dominant value dominates


@pytest.mark.parametrize("feature_type,density", [("bb", 1.0), ("bb", 0.5), ("dd", 1.0), ("gp", 0.5)])
def test_posterior_enum_through_tare_and_sparsify(tmp_path, tare_on_oracle, sparsify_on_oracle, posterior_enum_on_oracle,  # noqa: F811
                                                  feature_type, density):
    """loom's chi-square test on rows that went through loom_tare + loom_sparsify, as the reference's own test does by
    default (loom/test/test_posterior_enum.py:675-710): the sampled latents of the diff path follow the posterior"""
    import pe_util
    checked = tared = 0
    for i, (R, C, kinds) in enumerate([(3, 2, False), (4, 1, False), (3, 2, True)]):
        d = tmp_path / ("p%d" % i)
        d.mkdir()
        gof, info = posterior_enum_case(posterior_enum_on_oracle, d, R, C, feature_type, density, kinds, seed=30 + i,
                                        ingest=(tare_on_oracle, sparsify_on_oracle))
        tared += len(pb_schema.load_stream(d / "tares.pbs.gz", M()["ProductValue"]))
        assert gof is None or gof > pe_util.MIN_GOODNESS_OF_FIT, (R, C, gof, info)
        checked += gof is not None
    assert checked >= 1
    if feature_type in ("bb", "dd") and density == 1.0:
        assert tared >= 1          # a tare row was actually in play


def test_store_written_by_tasks_infer_is_served_by_loom_query(tmp_path, oracle_abi, query_on_oracle):  # noqa: F811
    """loom_b200.tasks.infer (all samples as chains on one engine) writing the store layout of src/store.hpp:68-89,
    config.pb.gz per sample included; loom_query then serves it"""
    from loom_b200 import hyperprior, tasks
    from test_z_loom_infer import write_inputs
    from test_query_server import expected_scores
    work = tmp_path / "work"
    work.mkdir()
    model, block, rowids, paths = write_inputs(work, n_rows=150, seed=6, tared=True)
    root = tmp_path / "store"
    (root / "ingest").mkdir(parents=True)
    shutil.copy(paths["rows"], root / "ingest" / "diffs.pbs.gz")
    shutil.copy(paths["tares"], root / "ingest" / "tares.pbs.gz")
    feature_sizes = [{"type": "dd", "dim": len(f["alphas"])} if f["type"] == "dd" else {"type": f["type"]} for f in model.features]
    inits, outs = [], []
    for s in range(N_SAMPLES):
        init = tasks.generate_init(feature_sizes, hyperprior.defaults(), seed=s)
        pbio.ModelFile.create(init, GRIDS).write(work / ("init.%d.pb.gz" % s))
        inits.append(str(work / ("init.%d.pb.gz" % s)))
        sd = root / "samples" / ("sample.%d" % s)
        sd.mkdir(parents=True)
        outs.append({"config": str(sd / "config.pb.gz"), "model": str(sd / "model.pb.gz"), "groups": str(sd / "groups"),
                     "assign": str(sd / "assign.pbs.gz")})
    config = {"schedule": {"extra_passes": 1.5, "small_data_size": 50.0}, "kernels": {"kind": {"iterations": 2, "empty_kind_count": 2}}}
    tasks.infer(paths["rows"], inits, outs, config, tares_in=paths["tares"], seeds=[11, 12, 13], abi=oracle_abi)
    assert [pbio.config_read(o["config"]).seed for o in outs] == [11, 12, 13]
    assert pbio.config_read(outs[0]["config"]).as_dict()["kernels"]["kind"]["iterations"] == 2
    qc = pbio.config_defaults()
    pbio.config_write(root / "query_config.pb.gz", qc)
    table = some_rows(model, block, range(5))
    requests = []
    for i, row in enumerate(table):
        r = new_request(i)
        fill_diff(r.score.data, model, cells_of(row))
        requests.append(r)
    responses = serve(query_on_oracle, root, requests, tmp_path)
    got = np.array([r.score.score for r in responses])
    assert np.allclose(got, expected_scores(oracle_abi, root, model, table), rtol=1e-5, atol=1e-5)


def test_loom_mix(tmp_path, infer_on_oracle, oracle_abi):  # noqa: F811
    """Loom::mix through the process (src/mix.cc): a fully assigned state, sample_skip sweeps with the kind and hyper
    kernels, dumped again -- every statistic of the dump recomputes from the rows and the dumped assignments, and the
    state moved"""
    import subprocess
    from test_query_server import build_on_oracle
    from test_z_loom_infer import check_outputs, write_inputs

    class Factory:                                   # build_on_oracle wants a tmp_path_factory
        @staticmethod
        def mktemp(name):
            d = tmp_path / name
            d.mkdir()
            return d
    exe = build_on_oracle(Factory, "loom_mix.cc", "loom_mix_oracle")
    model, block, rowids, paths = write_inputs(tmp_path, n_rows=200, seed=12, extra_passes=1.0, kind_iterations=2, hyper=True)
    first = run_infer(infer_on_oracle, paths, tmp_path, tag="first")
    c = pbio.config_read(paths["config"])
    c.generate_sample_skip, c.seed = 3, 99
    pbio.config_write(paths["config"], c)
    out = {"model": str(tmp_path / "mixed.model.pb.gz"), "groups": str(tmp_path / "mixed.groups"),
           "assign": str(tmp_path / "mixed.assign.pbs.gz")}
    proc = subprocess.run([exe, paths["config"], paths["rows"], first["model"], first["groups"], first["assign"],
                           out["model"], out["groups"], out["assign"]], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    check_outputs(oracle_abi, block, rowids, out)
    a0, a1 = pbio.assign_read(first["assign"]), pbio.assign_read(out["assign"])
    assert sorted(a0[0].tolist()) == sorted(a1[0].tolist())
    assert a0[1].shape != a1[1].shape or not np.array_equal(a0[1], a1[1])            # the chain moved
    c.generate_sample_skip = 0
    pbio.config_write(paths["config"], c)
    proc = subprocess.run([exe, paths["config"], paths["rows"], first["model"], first["groups"], first["assign"],
                           out["model"], out["groups"], out["assign"]], capture_output=True, timeout=600)
    assert proc.returncode != 0 and b"zero diversity" in proc.stderr                 # src/loom.cc:569


def test_loom_generate_then_mix(tmp_path, oracle_abi):
    """loom.generate's two steps through the processes (src/generate.cc, src/mix.cc): a dataset synthesised from a model
    -- rows, model, groups, assignments on disk -- whose every dumped statistic recomputes from the rows it wrote, then
    mixed; the config's row count and density are honoured"""
    import subprocess
    from test_query_server import build_on_oracle
    from test_z_loom_infer import TYPES, check_outputs

    class Factory:
        @staticmethod
        def mktemp(name):
            d = tmp_path / name
            d.mkdir(exist_ok=True)
            return d
    gen = build_on_oracle(Factory, "loom_generate.cc", "loom_generate_oracle")
    mix = build_on_oracle(Factory, "loom_mix.cc", "loom_mix_oracle")
    model, _, _ = synth.generate(10, TYPES, density=1.0, seed=2)
    pbio.ModelFile.create(model, GRIDS).write(tmp_path / "init.pb.gz")
    c = pbio.config_defaults()
    c.seed, c.generate_row_count, c.generate_density, c.generate_sample_skip = 21, 250, 0.6, 2
    c.kind_iterations, c.kind_empty_kind_count, c.hyper_run = 2, 2, 1
    pbio.config_write(tmp_path / "config.pb.gz", c)
    out = {"model": str(tmp_path / "gen.model.pb.gz"), "groups": str(tmp_path / "gen.groups"),
           "assign": str(tmp_path / "gen.assign.pbs.gz")}
    rows_path = tmp_path / "rows.pbs.gz"
    proc = subprocess.run([gen, str(tmp_path / "config.pb.gz"), str(tmp_path / "init.pb.gz"), str(rows_path),
                           out["model"], out["groups"], out["assign"]], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    mf = pbio.ModelFile.read(out["model"])
    rows = pbio.rows_read(rows_path, mf)
    assert rows.n_rows == 250 and rows.rowids.tolist() == list(range(250)) and not rows.row_has_tare.any()
    cat8, wide, mask = rows.to_columns(mf)
    block = RowBlock(250, cat8, wide, mask)
    assert abs(block.observed().mean() - 0.6) < 0.03
    check_outputs(oracle_abi, block, rows.rowids, out)
    mixed = {"model": str(tmp_path / "mix.model.pb.gz"), "groups": str(tmp_path / "mix.groups"),
             "assign": str(tmp_path / "mix.assign.pbs.gz")}
    proc = subprocess.run([mix, str(tmp_path / "config.pb.gz"), str(rows_path), out["model"], out["groups"], out["assign"],
                           mixed["model"], mixed["groups"], mixed["assign"]], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    check_outputs(oracle_abi, block, rows.rowids, mixed)
    # --none discards the state, the rows are still written
    proc = subprocess.run([gen, str(tmp_path / "config.pb.gz"), str(tmp_path / "init.pb.gz"), str(tmp_path / "rows2.pbs.gz"),
                           "--none", "--none", "--none"], capture_output=True, timeout=600)
    assert proc.returncode == 0 and pbio.stream_stats(tmp_path / "rows2.pbs.gz")[0] == 250
