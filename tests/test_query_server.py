"""loom_query -- the process loom.runner starts for loom.query (src/query.cc, src/query_server.cc) --
built against the oracle's lo_ ABI and driven the way loom drives it: a store directory with several
samples written by loom_infer, a stream of Query.Request in, a stream of Query.Response out (files, and
the piped stdin / stdout server).  Requests are written and responses parsed with google.protobuf, an
independent implementation of the wire format.  The GPU run of the product binary is in
tests/test_z_loom_infer.py."""
import itertools
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import pb_schema
from loom_b200 import Engine, RowBlock, pbio
from test_pbio import M, fill_value, io_lib  # noqa: F401  (io_lib: autouse fixture builds the codec)
from test_query import MIN_P, chi_square_p
from test_z_loom_infer import HOST, REPO, infer_on_oracle, run, write_inputs  # noqa: F401

N_SAMPLES = 3


def build_on_oracle(tmp_path_factory, source, name):
    exe = tmp_path_factory.mktemp("bin") / name
    oracle_dir, io_dir = os.path.join(REPO, "oracle"), os.path.join(REPO, "loom_b200", "io")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-DLB_NAME(x)=lo_##x",
                           os.path.join(HOST, source), "-o", str(exe),
                           "-L" + oracle_dir, "-lloom_oracle", "-L" + io_dir, "-lloomb200_io",
                           "-Wl,-rpath," + oracle_dir, "-Wl,-rpath," + io_dir])
    return str(exe)


@pytest.fixture(scope="module")
def query_on_oracle(tmp_path_factory, oracle_abi):
    return build_on_oracle(tmp_path_factory, "loom_query.cc", "loom_query_oracle")


def make_store(root, infer_exe, tared=False, n_rows=300):
    """ingest/diffs.pbs.gz (+ tares), samples/sample.N/{config.pb.gz, model.pb.gz, groups/} (src/store.hpp:68-89)"""
    work = root / "work"
    work.mkdir(parents=True)
    model, block, rowids, paths = write_inputs(work, n_rows=n_rows, seed=4, extra_passes=1.5, tared=tared)
    (root / "ingest").mkdir()
    shutil.copy(paths["rows"], root / "ingest" / "diffs.pbs.gz")
    if tared:
        shutil.copy(paths["tares"], root / "ingest" / "tares.pbs.gz")
    for s in range(N_SAMPLES):
        c = pbio.config_read(paths["config"])
        c.seed = 100 + s
        pbio.config_write(paths["config"], c)
        out = run(infer_exe, paths, work, tag="s%d" % s)
        d = root / "samples" / ("sample.%d" % s)
        d.mkdir(parents=True)
        shutil.copy(paths["config"], d / "config.pb.gz")
        shutil.copy(out["model"], d / "model.pb.gz")
        shutil.copytree(out["groups"], d / "groups")
    c = pbio.config_defaults()
    c.seed = 9
    pbio.config_write(root / "query_config.pb.gz", c)
    return model, block, rowids


@pytest.fixture(scope="module")
def store(tmp_path_factory, infer_on_oracle):  # noqa: F811
    root = tmp_path_factory.mktemp("store")
    model, block, rowids = make_store(root, infer_on_oracle)
    return root, model, block, rowids


def latents(abi, root, table_model, table):
    """the store's samples as engines holding `table` (what the server does per request)"""
    out = []
    for s in range(N_SAMPLES):
        d = root / "samples" / ("sample.%d" % s)
        mf = pbio.ModelFile.read(d / "model.pb.gz")
        model = mf.model()
        e = Engine(abi)
        e.set_model(model)
        e.set_rows(RowBlock.from_table(model, table))
        for k in range(model.n_kinds):
            counts, ints, flts = pbio.groups_read(pbio.mixture_path(d / "groups", k), mf, model.kind_features(k))
            e.set_groups(k, len(counts), counts, ints, flts)
        out.append(e)
    return out


def expected_scores(abi, root, model, table):
    per = np.stack([e.query_score().astype(np.float64) for e in latents(abi, root, model, table)])
    m = per.max(axis=0)
    return m + np.log(np.exp(per - m).sum(axis=0)) - np.log(N_SAMPLES)


def cells_of(row):
    return {f: v for f, v in enumerate(row) if v is not None}


def new_request(i):
    r = M()["Query.Request"]()
    r.id = "request-%d" % i
    return r


def fill_diff(diff, model, cells, form="auto"):
    fill_value(diff.pos, model, cells, form)
    diff.neg.observed.sparsity = 0


def fill_observed(obs, model, feats, form):
    if form == "dense":
        obs.sparsity = 2
        obs.dense.extend([f in feats for f in range(model.n_features)])
    elif form == "sparse":
        obs.sparsity = 1
        obs.sparse.extend(sorted(feats))
    else:
        assert not feats
        obs.sparsity = 0


def serve(exe, root, requests, tmp):
    pb_schema.dump_stream(tmp / "requests.pbs.gz", requests)
    proc = subprocess.run([exe, str(root), str(tmp / "requests.pbs.gz"), str(root / "query_config.pb.gz"),
                           str(tmp / "responses.pbs.gz"), "--none"], capture_output=True, timeout=600)
    assert proc.returncode == 0, proc.stderr.decode()
    responses = pb_schema.load_stream(tmp / "responses.pbs.gz", M()["Query.Response"])
    assert [r.id for r in responses] == [r.id for r in requests]
    assert all(r.IsInitialized() for r in responses)
    return responses


def value_cells(model, pv):
    """{feature: value} of a ProductValue (src/product_value.hpp:586-767)"""
    F = model.n_features
    sp = pv.observed.sparsity
    feats = (list(range(F)) if sp == 3 else [f for f in range(F) if pv.observed.dense[f]] if sp == 2
             else list(pv.observed.sparse) if sp == 1 else [])
    b, c, r = iter(pv.booleans), iter(pv.counts), iter(pv.reals)
    out = {}
    for f in feats:
        t = model.features[f]["type"]
        out[f] = int(next(b)) if t == "bb" else float(next(r)) if t == "nich" else int(next(c))
    assert not list(b) and not list(c) and not list(r)
    return out


def some_rows(model, block, rows):
    obs = block.observed()
    table = []
    for r in rows:
        row = [None] * model.n_features
        for f in range(model.n_features):
            if obs[f, r]:
                if f < model.n8():
                    row[f] = int(block.cat8[f, r])
                elif model.features[f]["type"] == "nich":
                    row[f] = float(block.wide[f - model.n8(), r:r + 1].view(np.float32)[0])
                else:
                    row[f] = int(block.wide[f - model.n8(), r])
        table.append(row)
    return table


def test_score(store, query_on_oracle, oracle_abi, tmp_path):
    root, model, block, rowids = store
    table = some_rows(model, block, range(12)) + [[None] * model.n_features]
    requests = []
    for i, row in enumerate(table):
        r = new_request(i)
        fill_diff(r.score.data, model, cells_of(row), ["auto", "dense", "sparse"][i % 3] if cells_of(row) else "auto")
        requests.append(r)
    responses = serve(query_on_oracle, root, requests, tmp_path)
    got = np.array([r.score.score for r in responses])
    want = expected_scores(oracle_abi, root, model, table)
    assert all(r.HasField("score") and not r.error and not r.HasField("sample") for r in responses)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)
    assert abs(got[-1]) < 1e-5                       # the empty row has probability one


def test_sample_marginals_match_the_servers_own_scores(store, query_on_oracle, tmp_path):
    """test_query_math.py:65-92: p(value | conditional) from Score against the frequencies of Sample, over the
    MIXTURE of latents"""
    root, model, block, rowids = store
    F = model.n_features
    cond = cells_of(some_rows(model, block, [0])[0])
    discrete = [f for f in range(F) if model.features[f]["type"] in ("bb", "dd") and
                (model.features[f]["type"] == "bb" or len(model.features[f]["alphas"]) <= 16)]
    to_sample = [f for f in discrete if f % 2 == 0] + [F - 1]      # a real-valued feature rides along
    for f in to_sample:
        cond.pop(f, None)
    n = 4000
    r = new_request(0)
    fill_diff(r.sample.data, model, cond)
    fill_observed(r.sample.to_sample, model, to_sample, "dense")
    r.sample.sample_count = n
    requests = [r]
    base = new_request(1)
    fill_diff(base.score.data, model, cond)
    requests.append(base)
    candidates = []
    for f in to_sample[:-1]:
        dim = 2 if model.features[f]["type"] == "bb" else len(model.features[f]["alphas"])
        for v in range(dim):
            q = new_request(len(requests))
            fill_diff(q.score.data, model, {**cond, f: v})
            requests.append(q)
            candidates.append((f, v))
    responses = serve(query_on_oracle, root, requests, tmp_path)
    samples = [value_cells(model, s.pos) for s in responses[0].sample.samples]
    assert len(samples) == n and all(sorted(s) == sorted(to_sample) for s in samples)
    assert all(s.neg.observed.sparsity == 0 for s in responses[0].sample.samples)
    assert np.isfinite([s[F - 1] for s in samples]).all()
    base_score = responses[1].score.score
    p = {fv: np.exp(resp.score.score - base_score) for fv, resp in zip(candidates, responses[2:])}
    for f in to_sample[:-1]:
        values = sorted(v for (g, v) in p if g == f)
        probs = np.array([p[f, v] for v in values])
        assert abs(probs.sum() - 1) < 1e-4
        counts = [sum(1 for s in samples if s[f] == v) for v in values]
        assert chi_square_p(counts, probs) > MIN_P, f


def test_entropy_matches_enumeration(store, query_on_oracle, oracle_abi, tmp_path):
    """H(S | conditional) of small discrete sets by enumerating every value combination through Score"""
    root, model, block, rowids = store
    F = model.n_features
    bbs = [f for f in range(F) if model.features[f]["type"] == "bb"]
    dds = [f for f in range(F) if model.features[f]["type"] == "dd" and len(model.features[f]["alphas"]) == 16]
    a, b, c = bbs[0], bbs[1], dds[0]
    cond = cells_of(some_rows(model, block, [3])[0])
    for f in (a, b, c):
        cond.pop(f, None)
    row_sets, col_sets = [set(), {a}], [set(), {b}, {b, c}]
    r = new_request(0)
    for s in row_sets:
        fill_observed(r.entropy.row_sets.add(), model, s, "sparse" if s else "none")
    for s in col_sets:
        fill_observed(r.entropy.col_sets.add(), model, s, "dense" if s else "none")
    fill_diff(r.entropy.conditional, model, cond)
    n = 3000
    r.entropy.sample_count = n
    resp = serve(query_on_oracle, root, [r], tmp_path)[0]
    assert not resp.error and len(resp.entropy.means) == 6 and len(resp.entropy.variances) == 6

    def dim(f):
        return 2 if model.features[f]["type"] == "bb" else len(model.features[f]["alphas"])

    k = 0
    for rs in row_sets:
        for cs in col_sets:
            feats = sorted(rs | cs)
            combos = list(itertools.product(*[range(dim(f)) for f in feats]))
            table = [[cond.get(f) for f in range(F)]]
            for combo in combos:
                row = list(table[0])
                for f, v in zip(feats, combo):
                    row[f] = v
                table.append(row)
            s = expected_scores(oracle_abi, root, model, table)
            logp = s[1:] - s[0]
            assert abs(np.exp(logp).sum() - 1) < 1e-5
            exact = -(np.exp(logp) * logp).sum()
            mean, var = resp.entropy.means[k], resp.entropy.variances[k]
            if not feats:
                assert abs(mean) < 1e-5 and var < 1e-9
            else:
                assert var > 0 and abs(var * n - ((np.exp(logp) * logp ** 2).sum() - exact ** 2)) < 0.25 * var * n
                assert abs(mean - exact) < 4.5 * np.sqrt(var), (feats, mean, exact, var)
            k += 1


def test_entropy_errors(store, query_on_oracle, tmp_path):
    root, model, block, rowids = store
    r0 = new_request(0)
    fill_observed(r0.entropy.row_sets.add(), model, {0}, "sparse")
    fill_observed(r0.entropy.col_sets.add(), model, {1}, "sparse")
    fill_diff(r0.entropy.conditional, model, {})
    r0.entropy.sample_count = 1
    r1 = new_request(1)
    bad = r1.entropy.row_sets.add()
    bad.sparsity = 1
    bad.sparse.extend([3, 2])
    fill_diff(r1.entropy.conditional, model, {})
    r1.entropy.sample_count = 10
    r2 = new_request(2)
    fill_observed(r2.entropy.row_sets.add(), model, {0}, "sparse")
    fill_observed(r2.entropy.col_sets.add(), model, {1}, "sparse")
    fill_diff(r2.entropy.conditional, model, {0: 1})
    r2.entropy.sample_count = 10
    responses = serve(query_on_oracle, root, [r0, r1, r2], tmp_path)
    assert list(responses[0].error) == ["invalid request.entropy.sample_count"]       # query_server.cc:259-262
    assert list(responses[1].error) == ["invalid request.entropy.row_sets"]           # :247-252
    assert "overlaps the conditional" in responses[2].error[0]
    assert not any(r.HasField("entropy") for r in responses)


def test_invalid_data_is_reported_and_the_server_carries_on(store, query_on_oracle, tmp_path):
    root, model, block, rowids = store
    dd = next(f for f in range(model.n_features) if model.features[f]["type"] == "dd")
    r0 = new_request(0)
    fill_diff(r0.score.data, model, {dd: 10 ** 6})            # out of range for the feature
    fill_diff(r0.sample.data, model, {})
    r0.sample.to_sample.sparsity = 2                           # dense of the wrong size
    r0.sample.to_sample.dense.extend([True])
    r0.sample.sample_count = 3
    r1 = new_request(1)
    fill_diff(r1.score.data, model, {dd: 1})
    fill_diff(r1.sample.data, model, {dd: 1})
    fill_observed(r1.sample.to_sample, model, {0}, "sparse")
    r1.sample.sample_count = 3
    responses = serve(query_on_oracle, root, [r0, r1], tmp_path)
    assert sorted(responses[0].error) == ["invalid request.sample.to_sample", "invalid request.score.data"]
    assert not responses[0].HasField("score") and not responses[0].HasField("sample")
    assert not responses[1].error and responses[1].HasField("score") and len(responses[1].sample.samples) == 3


def test_tare_ids_are_validated(store, query_on_oracle, tmp_path):
    """src/query_server.cc:77-82, 178-183: a tare id the store does not hold (this store has no tare row at all)"""
    root, model, block, rowids = store
    r0 = new_request(0)
    fill_diff(r0.score.data, model, {0: 1})
    r0.score.data.tares.append(0)
    fill_diff(r0.sample.data, model, {})
    r0.sample.data.tares.append(3)
    fill_observed(r0.sample.to_sample, model, {1}, "sparse")
    r0.sample.sample_count = 2
    fill_diff(r0.score_derivative.update_data, model, {0: 1})
    r0.score_derivative.update_data.tares.append(0)
    r0.score_derivative.row_limit = 1
    r1 = new_request(1)
    fill_diff(r1.score.data, model, {0: 1})
    responses = serve(query_on_oracle, root, [r0, r1], tmp_path)
    assert sorted(responses[0].error) == ["invalid request.sample.data.tares", "invalid request.score.data.tares",
                                          "invalid request.score_derivative.update_data"]
    assert not responses[0].HasField("score") and not responses[0].HasField("sample")
    assert not responses[0].HasField("score_derivative")
    assert responses[1].HasField("score") and not responses[1].error


def test_score_derivative(store, query_on_oracle, tmp_path):
    root, model, block, rowids = store
    table = some_rows(model, block, range(8))
    update = cells_of(table[0])
    r0 = new_request(0)
    fill_diff(r0.score_derivative.update_data, model, update)
    for row in table[:5]:
        fill_diff(r0.score_derivative.score_data.add(), model, cells_of(row))
    r0.score_derivative.row_limit = 3
    r1 = new_request(1)                                        # no score_data: every row of ingest/diffs.pbs.gz
    fill_diff(r1.score_derivative.update_data, model, update)
    r1.score_derivative.row_limit = 1000
    r2 = new_request(2)                                        # the latents are back to their state: same scores as before
    fill_diff(r2.score.data, model, update)
    before = new_request(3)
    fill_diff(before.score.data, model, update)
    responses = serve(query_on_oracle, root, [before, r0, r1, r2], tmp_path)
    d0, d1 = responses[1].score_derivative, responses[2].score_derivative
    assert len(d0.ids) == 3 and len(d0.score_diffs) == 3 and set(d0.ids) <= set(range(5))
    assert list(d0.score_diffs) == sorted(d0.score_diffs, reverse=True)
    assert d0.ids[0] == 0 and d0.score_diffs[0] > 0            # a second copy of the row makes that row more likely
    assert len(d1.ids) == block.n_rows and sorted(d1.ids) == sorted(rowids.tolist())
    assert list(d1.score_diffs) == sorted(d1.score_diffs, reverse=True)
    assert np.isfinite(d1.score_diffs).all() and d1.ids[0] == rowids[0]
    assert responses[3].score.score == pytest.approx(responses[0].score.score, abs=1e-5)


def mixed_session_case(exe, root, model, block, oracle_abi, tmp_path):
    """one session mixing every call, several sub-requests per message (also run on the device by
    tests/test_zz_query_gpu.py): Score against the oracle's latents, Sample against the server's own scores,
    Entropy against enumeration, ScoreDerivative leaves the latents as they were"""
    table = some_rows(model, block, range(6))
    requests = []
    for i, row in enumerate(table):
        r = new_request(i)
        fill_diff(r.score.data, model, cells_of(row))
        requests.append(r)
    bb = next(f for f in range(model.n_features) if model.features[f]["type"] == "bb")
    cond = {f: v for f, v in cells_of(table[0]).items() if f != bb}
    n = 4000
    s = new_request(len(requests))
    fill_diff(s.sample.data, model, cond)
    fill_observed(s.sample.to_sample, model, {bb}, "sparse")
    s.sample.sample_count = n
    fill_diff(s.score.data, model, cond)
    requests.append(s)
    for v in (0, 1):
        r = new_request(len(requests))
        fill_diff(r.score.data, model, {**cond, bb: v})
        requests.append(r)
    d = new_request(len(requests))
    fill_diff(d.score_derivative.update_data, model, cells_of(table[0]))
    for row in table[:3]:
        fill_diff(d.score_derivative.score_data.add(), model, cells_of(row))
    d.score_derivative.row_limit = 2
    fill_observed(d.entropy.row_sets.add(), model, {bb}, "sparse")
    fill_observed(d.entropy.col_sets.add(), model, set(), "none")
    fill_diff(d.entropy.conditional, model, cond)
    d.entropy.sample_count = 500
    requests.append(d)
    again = new_request(len(requests))
    fill_diff(again.score.data, model, cells_of(table[0]))
    requests.append(again)
    responses = serve(exe, root, requests, tmp_path)
    assert not any(r.error for r in responses)
    got = np.array([r.score.score for r in responses[:6]])
    want = expected_scores(oracle_abi, root, model, table)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)
    samples = [value_cells(model, x.pos) for x in responses[6].sample.samples]
    assert len(samples) == n and all(list(x) == [bb] for x in samples)
    p = np.exp(np.array([responses[7].score.score, responses[8].score.score]) - responses[6].score.score)
    assert abs(p.sum() - 1) < 1e-4
    ones = sum(x[bb] for x in samples)
    assert chi_square_p([n - ones, ones], p) > MIN_P
    sd = responses[9].score_derivative
    assert len(sd.ids) == 2 and sd.ids[0] == 0 and sd.score_diffs[0] > 0 and sd.score_diffs[0] >= sd.score_diffs[1]
    exact = -(p * np.log(p)).sum()
    mean, var = responses[9].entropy.means[0], responses[9].entropy.variances[0]
    assert abs(mean - exact) < 4.5 * np.sqrt(var) + 1e-6
    assert responses[10].score.score == pytest.approx(responses[0].score.score, abs=1e-5)


def test_mixed_session(store, query_on_oracle, oracle_abi, tmp_path):
    root, model, block, rowids = store
    mixed_session_case(query_on_oracle, root, model, block, oracle_abi, tmp_path)


def test_tared_store(tmp_path, infer_on_oracle, query_on_oracle, oracle_abi):  # noqa: F811
    """rows referencing the tare (Diff{pos, neg, tares:[0]}): the score of a diff equals the score of the row it encodes"""
    from loom_b200 import tare
    root = tmp_path / "store"
    model, block, rowids = make_store(root, infer_on_oracle, tared=True, n_rows=200)
    tare_obs, tare_val = tare.make_tare(model, block)
    diff = tare.sparsify(model, block, tare_obs, tare_val)
    requests = []
    rows = list(range(6))
    for i in rows:
        r = new_request(i)
        d = r.score.data
        p0, p1 = diff["pos_ptr"][i], diff["pos_ptr"][i + 1]
        n0, n1 = diff["neg_ptr"][i], diff["neg_ptr"][i + 1]
        pos = {}
        for f, v in zip(diff["pos_feature"][p0:p1], diff["pos_value"][p0:p1]):
            t = model.features[int(f)]["type"]
            pos[int(f)] = float(np.uint32(v).view(np.float32)) if t == "nich" else int(v)
        fill_value(d.pos, model, pos, "auto")
        fill_value(d.neg, model, {int(f): 0 for f in diff["neg_feature"][n0:n1]}, "auto")
        d.tares.append(0)
        requests.append(r)
    e = new_request(len(rows))
    fill_diff(e.entropy.conditional, model, {})
    e.entropy.conditional.tares.append(0)                      # conditional = the tare row itself
    free = [f for f in range(model.n_features) if not tare_obs[f]][:1]
    fill_observed(e.entropy.row_sets.add(), model, set(free), "sparse" if free else "none")
    fill_observed(e.entropy.col_sets.add(), model, set(), "none")
    e.entropy.sample_count = 50
    requests.append(e)
    # a diff that does not fit the tare it references: neg names a cell the tare lacks / pos sets a cell the tare holds
    free_cell = next(f for f in range(model.n_features) if not tare_obs[f] and model.features[f]["type"] == "bb")
    held_cell = next(f for f in range(model.n_features) if tare_obs[f])
    for bad_pos, bad_neg in (({}, {free_cell: 0}), ({held_cell: int(tare_val[held_cell])}, {})):
        r = new_request(len(requests))
        fill_value(r.score.data.pos, model, bad_pos, "sparse" if bad_pos else "auto")
        fill_value(r.score.data.neg, model, bad_neg, "sparse" if bad_neg else "auto")
        r.score.data.tares.append(0)
        requests.append(r)
    responses = serve(query_on_oracle, root, requests, tmp_path)
    assert [list(r.error) for r in responses[-2:]] == [["invalid request.score.data"]] * 2
    assert not any(r.HasField("score") for r in responses[-2:])
    responses = responses[:-2]
    got = np.array([r.score.score for r in responses[:len(rows)]])
    want = expected_scores(oracle_abi, root, model, some_rows(model, block, rows))
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)
    assert not responses[-1].error and len(responses[-1].entropy.means) == 1


def test_piped_server_answers_each_request_before_the_next(store, query_on_oracle):
    """loom.query's piped server (loom/runner.py:376-381): `-` `-`; the client sends a request and WAITS for its
    response, so the server may not sit on a read buffer"""
    root, model, block, rowids = store
    proc = subprocess.Popen([query_on_oracle, str(root), "-", str(root / "query_config.pb.gz"), "-", "--none"],
                            stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        scores = []
        for i in range(3):
            r = new_request(i)
            fill_diff(r.score.data, model, cells_of(some_rows(model, block, [i])[0]))
            data = r.SerializeToString()
            proc.stdin.write(struct.pack("<I", len(data)) + data)
            proc.stdin.flush()
            (n,) = struct.unpack("<I", proc.stdout.read(4))     # blocks until the server answers
            resp = M()["Query.Response"]()
            resp.ParseFromString(proc.stdout.read(n))
            assert resp.id == r.id and resp.HasField("score")
            scores.append(resp.score.score)
        proc.stdin.close()
        assert proc.wait(timeout=60) == 0
        assert len(set(scores)) == 3
    finally:
        proc.kill()
        proc.wait()


def test_missing_samples_abort_with_looms_message(tmp_path, query_on_oracle):
    (tmp_path / "ingest").mkdir()
    c = pbio.config_defaults()
    pbio.config_write(tmp_path / "config.pb.gz", c)
    proc = subprocess.run([query_on_oracle, str(tmp_path), "--none", str(tmp_path / "config.pb.gz"), "--none", "--none"],
                          capture_output=True, timeout=60)
    assert proc.returncode != 0 and b"no samples were found at" in proc.stderr     # src/multi_loom.cc:69


def test_seed(store, query_on_oracle, tmp_path):
    """loom/test/test_query.py:207-227: the same config seed gives the same responses, another seed different ones"""
    root, model, block, rowids = store
    requests = []
    for i in range(4):
        r = new_request(i)
        fill_diff(r.sample.data, model, cells_of(some_rows(model, block, [i])[0]) if i % 2 else {})
        to_sample = {f for f in range(model.n_features) if f % 2 == i % 2}
        if i % 2:
            to_sample -= set(cells_of(some_rows(model, block, [i])[0]))
        fill_observed(r.sample.to_sample, model, to_sample, "dense")
        r.sample.sample_count = 5
        fill_diff(r.score.data, model, {})
        requests.append(r)
    pb_schema.dump_stream(tmp_path / "requests.pbs.gz", requests)
    outs = []
    for tag, seed in (("a", 0), ("b", 0), ("c", 10)):
        c = pbio.config_defaults()
        c.seed = seed
        pbio.config_write(tmp_path / (tag + ".config.pb.gz"), c)
        proc = subprocess.run([query_on_oracle, str(root), str(tmp_path / "requests.pbs.gz"), str(tmp_path / (tag + ".config.pb.gz")),
                               str(tmp_path / (tag + ".responses.pbs")), "--none"], capture_output=True, timeout=600)
        assert proc.returncode == 0, proc.stderr.decode()
        outs.append(open(tmp_path / (tag + ".responses.pbs"), "rb").read())
    assert outs[0] == outs[1] and outs[0] != outs[2]
    assert len(pb_schema.load_stream(tmp_path / "a.responses.pbs", M()["Query.Response"])) == 4


def test_hostile_but_well_formed_requests_are_answered(store, query_on_oracle, tmp_path):
    """values a fuzzer found the server aborting on: a neg part without a tare reference (ignored, as the reference scores
    such a diff from pos alone, src/query_server.cc:121-125), non-finite reals, zero counts and limits"""
    root, model, block, rowids = store
    F = model.n_features
    nich = next(f for f in range(F) if model.features[f]["type"] == "nich")
    gp = next(f for f in range(F) if model.features[f]["type"] == "gp")
    r0 = new_request(0)
    fill_value(r0.score.data.pos, model, {0: 1}, "sparse")
    fill_value(r0.score.data.neg, model, {1: 0, 2: 1}, "sparse")          # no tares: neg is ignored
    r1 = new_request(1)
    fill_diff(r1.score.data, model, {0: 1})
    r2 = new_request(2)
    fill_diff(r2.score.data, model, {nich: float("nan"), gp: 10 ** 6})
    fill_diff(r2.sample.data, model, {nich: float("inf")})
    fill_observed(r2.sample.to_sample, model, {0}, "sparse")
    r2.sample.sample_count = 0
    fill_diff(r2.score_derivative.update_data, model, {0: 1})
    r2.score_derivative.row_limit = 0
    fill_diff(r2.score_derivative.score_data.add(), model, {0: 0})
    responses = serve(query_on_oracle, root, [r0, r1, r2], tmp_path)
    assert not any(r.error for r in responses)
    assert responses[0].score.score == pytest.approx(responses[1].score.score, abs=1e-6)
    assert len(responses[2].sample.samples) == 0 and len(responses[2].score_derivative.ids) == 0
