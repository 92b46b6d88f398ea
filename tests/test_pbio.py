"""loom's protobuf files through the native codec (loom_b200/io/lb_io.cc), checked against
google.protobuf as an independent implementation of the wire format (tests/pb_schema.py).

Anchors in the reference: src/schema.proto (field numbers; compared below when the tree is
present), src/protobuf_stream.hpp (framing), src/product_value.hpp:586-767 (sparsity forms),
src/cross_cat.cc:50-135 (model), src/product_mixture.cc:739-856 (groups),
src/assignments.cc:50-100 (assignments)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import pb_schema
from loom_b200 import _abi, pbio, synth, tare
from loom_b200.engine import Model, RowBlock

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SCHEMA = "/root/reference/src/schema.proto"


@pytest.fixture(scope="module", autouse=True)
def io_lib():
    d = os.path.join(REPO, "loom_b200", "io")
    srcs = [os.path.join(d, f) for f in ("lb_io.cc", "pbwire.hpp")] + [os.path.join(REPO, "include", "loom_b200_io.h")]
    if not os.path.exists(pbio.IO_LIB) or any(os.path.getmtime(s) > os.path.getmtime(pbio.IO_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    return pbio.lib()


M = pb_schema.messages


# ------------------------------------------------------------------------------------------ schema / ABI
@pytest.mark.skipif(not os.path.exists(REFERENCE_SCHEMA), reason="reference tree not present")
def test_wire_schema_matches_reference():
    """every message, field name, label, type and number we restate equals src/schema.proto's"""
    ref_msgs, ref_enums = pb_schema.flatten(pb_schema.parse(open(REFERENCE_SCHEMA).read()))
    our_msgs, our_enums = pb_schema.flatten(
        pb_schema.parse(open(os.path.join(pb_schema.GOLDEN, "loom_wire.proto")).read()))
    assert len(our_msgs) >= 35
    for name, fields in our_msgs.items():
        assert ref_msgs.get(name) == fields, name
    for name, values in our_enums.items():
        assert ref_enums.get(name) == values, name


def test_io_header_and_binding_agree(io_lib):
    text = open(os.path.join(REPO, "include", "loom_b200_io.h")).read()
    declared = sorted(set(re.findall(r"\blbio_(\w+)\s*\(", text)))
    assert declared == sorted(pbio.SIGNATURES)
    for name in declared:
        assert hasattr(io_lib, "lbio_" + name)


# ------------------------------------------------------------------------------------------ config / checkpoint
def test_config_defaults_and_round_trip(tmp_path):
    c = pbio.config_defaults()
    d = c.as_dict()
    assert d["schedule"]["extra_passes"] == 500.0 and d["kernels"]["kind"]["empty_kind_count"] == 32   # loom/config.py:34-74
    c.seed, c.extra_passes, c.kind_iterations, c.hyper_run = 1234567890123, 1.5, 7, 0
    for name in ("config.pb", "config.pb.gz"):
        path = tmp_path / name
        pbio.config_write(path, c)
        g = pb_schema.load_message(path, M()["Config"])
        assert g.IsInitialized()
        assert g.seed == 1234567890123 and g.schedule.extra_passes == 1.5 and g.kernels.kind.iterations == 7
        assert g.kernels.hyper.run is False and g.kernels.cat.empty_group_count == 1
        assert g.schedule.max_reject_iters == 100 and g.posterior_enum.sample_count == 100
        # and back: a file written by google.protobuf
        g.kernels.cat.empty_group_count = 3
        g.schedule.big_data_size = 5e6
        pb_schema.dump_message(path, g)
        back = pbio.config_read(path)
        assert back.cat_empty_group_count == 3 and back.big_data_size == 5e6 and back.seed == c.seed
        assert back.kind_iterations == 7 and back.hyper_run == 0 and back.has_query == 1
        # byte-identical with google's canonical serialisation
        pbio.config_write(path, back)
        raw = pb_schema._open(path, "rb").read()
        assert raw == g.SerializeToString()


def test_checkpoint_round_trip(tmp_path):
    g = M()["Checkpoint"]()
    g.finished, g.seed, g.tardis_iter, g.row_count = False, 2 ** 40 + 3, 17, 1000
    g.schedule.annealing_state, g.schedule.row_count, g.schedule.reject_iters = -123.456, 321, 4
    g.rows.unassigned_pos, g.rows.assigned_pos = 650, 329
    path = tmp_path / "checkpoint.pb.gz"
    pb_schema.dump_message(path, g)
    c = pbio.checkpoint_read(path)
    assert (c.finished, c.seed, c.tardis_iter, c.row_count) == (0, 2 ** 40 + 3, 17, 1000)
    assert (c.annealing_state, c.schedule_row_count, c.reject_iters) == (-123.456, 321, 4)
    assert (c.unassigned_pos, c.assigned_pos) == (650, 329)
    c.finished, c.tardis_iter = 1, 18
    pbio.checkpoint_write(path, c)
    g2 = pb_schema.load_message(path, M()["Checkpoint"])
    assert g2.IsInitialized() and g2.finished and g2.tardis_iter == 18 and g2.schedule.annealing_state == -123.456
    g.finished, g.tardis_iter = True, 18
    assert g2.SerializeToString() == g.SerializeToString()


# ------------------------------------------------------------------------------------------ model
TYPES = [("bb", 3), ("dd16", 2), ("dd256", 2), ("dpd5", 2), ("gp", 2), ("nich", 3)]
DPD_DICTS = [[7, 100, 3, 2 ** 20, 55], [0, 1, 2, 3, 4]]


def mixed_model(seed=0, n_kinds=3):
    rng = np.random.default_rng(seed)
    feats = synth.make_features(TYPES)
    for f in feats:     # distinct hypers so a swap would show
        for k in ("alpha", "beta", "inv_beta", "mu", "kappa", "sigmasq", "nu", "gamma"):
            if k in f:
                f[k] = float(np.float32(f[k] + rng.random()))
        if "alphas" in f:
            f["alphas"] = [float(np.float32(a + rng.random())) for a in f["alphas"]]
    f2k = rng.integers(0, n_kinds, len(feats)).astype(np.uint32)
    f2k[:n_kinds] = np.arange(n_kinds)
    kind_py = [(float(np.float32(1 + k)), float(np.float32(0.1 * k))) for k in range(n_kinds)]
    return Model(feats, f2k, kind_py, (np.float32(1.5), np.float32(0.25)))


def google_model(model, grids=None, dpd_dicts=DPD_DICTS):
    """the CrossCat message a loom front end would write for `model` (loom/generate.py:99-135)"""
    msg = M()["CrossCat"]()
    dpd_seen = 0
    dpd_of = {}
    for f, feat in enumerate(model.features):
        if feat["type"] == "dpd":
            dpd_of[f] = dpd_seen
            dpd_seen += 1
    for k in range(model.n_kinds):
        kind = msg.kinds.add()
        pm = kind.product_model
        pm.clustering.alpha, pm.clustering.d = model.kind_py[k]
        for f in model.kind_features(k):
            feat = model.features[f]
            kind.featureids.append(f)
            t = feat["type"]
            if t == "bb":
                s = pm.bb.add()
                s.alpha, s.beta = feat["alpha"], feat["beta"]
            elif t == "dd":
                pm.dd.add().alphas.extend(feat["alphas"])
            elif t == "dpd":
                s = pm.dpd.add()
                s.gamma, s.alpha = feat["gamma"], feat["alpha"]
                s.values.extend(dpd_dicts[dpd_of[f]])
                s.betas.extend(feat["betas"])
                s.counts.extend([0] * len(feat["betas"]))
            elif t == "gp":
                s = pm.gp.add()
                s.alpha, s.inv_beta = feat["alpha"], feat["inv_beta"]
            else:
                s = pm.nich.add()
                s.mu, s.kappa, s.sigmasq, s.nu = feat["mu"], feat["kappa"], feat["sigmasq"], feat["nu"]
    msg.topology.alpha, msg.topology.d = model.topology
    if grids:
        hp = msg.hyper_prior
        for name in ("topology", "clustering"):
            for a, d in grids.get(name, []):
                p = getattr(hp, name).add()
                p.alpha, p.d = a, d
        for t, params in pbio._GRID_KEYS.items():
            for p in params:
                getattr(getattr(hp, t), p).extend(grids.get(t, {}).get(p, []))
    return msg


GRIDS = {"topology": [(0.5, 0.0), (1.0, 0.25)], "clustering": [(2.0, 0.1), (4.0, 0.5), (8.0, 0.0)],
         "bb": {"alpha": [0.5, 1.0], "beta": [0.25, 2.0, 4.0]}, "dd": {"alpha": [0.5, 1.5]},
         "dpd": {"gamma": [1.0], "alpha": [0.5, 2.0]}, "gp": {"alpha": [1.0, 2.0], "inv_beta": [0.5]},
         "nich": {"mu": [-1.0, 1.0], "kappa": [0.5], "sigmasq": [1.0, 2.0], "nu": [2.0, 4.0, 8.0]}}


def assert_same_model(a, b):
    assert a.n_kinds == b.n_kinds and np.array_equal(a.f2k, b.f2k)
    assert np.array_equal(a.kind_py, b.kind_py) and np.array_equal(a.topology, b.topology)
    assert len(a.features) == len(b.features)
    for fa, fb in zip(a.features, b.features):
        assert fa["type"] == fb["type"]
        for key in fa:
            if key == "beta0":
                assert abs(fa[key] - fb[key]) < 1e-6
            elif key != "type":
                assert np.allclose(np.float32(fa[key]), np.float32(fb[key]), rtol=0, atol=0), key


def test_model_read_matches_google(tmp_path):
    model = mixed_model()
    msg = google_model(model, GRIDS)
    assert msg.IsInitialized()
    for name in ("model.pb.gz", "model.pb"):
        path = tmp_path / name
        pb_schema.dump_message(path, msg)
        mf = pbio.ModelFile.read(path)
        assert_same_model(mf.model(), model)
        got = mf.hyper_prior()
        for key, want in GRIDS.items():
            if isinstance(want, dict):
                for p, vals in want.items():
                    assert np.allclose(got[key][p], vals)
            else:
                assert np.allclose(got[key], want)
        assert [d.tolist() for d in mf.dpd_values()] == DPD_DICTS
        # write back: google.protobuf must parse the same message
        out = tmp_path / ("out." + name)
        mf.write(out)
        back = pb_schema.load_message(out, M()["CrossCat"])
        assert back.IsInitialized()
        assert back == msg
        assert back.SerializeToString() == msg.SerializeToString()


def test_model_create_from_engine_model(tmp_path):
    model = mixed_model(seed=3, n_kinds=2)
    mf = pbio.ModelFile.create(model, GRIDS, DPD_DICTS)
    path = tmp_path / "model.pb.gz"
    mf.write(path)
    assert pb_schema.load_message(path, M()["CrossCat"]) == google_model(model, GRIDS)
    again = pbio.ModelFile.read(path)
    assert_same_model(again.model(), model)
    # no hyper prior, identity dictionaries
    plain = pbio.ModelFile.create(model)
    assert plain.hyper_prior() is None and [d.tolist() for d in plain.dpd_values()] == [list(range(5))] * 2


def test_model_errors(tmp_path):
    model = mixed_model()
    msg = google_model(model)
    path = tmp_path / "bad.pb"
    # a feature in two kinds
    dup = M()["CrossCat"]()
    dup.CopyFrom(msg)
    dup.kinds[1].featureids[0] = dup.kinds[0].featureids[0]
    pb_schema.dump_message(path, dup)
    with pytest.raises(_abi.LoomError, match="duplicate feature|not in loom's order"):
        pbio.ModelFile.read(path)
    # BNB is not supported
    bnb = M()["CrossCat"]()
    bnb.CopyFrom(msg)
    s = bnb.kinds[0].product_model.bnb.add()
    s.alpha, s.beta, s.r = 1, 1, 1
    pb_schema.dump_message(path, bnb)
    with pytest.raises(_abi.LoomError, match="BetaNegativeBinomial"):
        pbio.ModelFile.read(path)
    with pytest.raises(_abi.LoomError, match="failed to open"):
        pbio.ModelFile.read(tmp_path / "missing.pb.gz")
    # truncated message
    raw = msg.SerializeToString()
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(_abi.LoomError):
        pbio.ModelFile.read(path)


# ------------------------------------------------------------------------------------------ rows
def random_table(model, n_rows, rng, density=0.6, dpd_dicts=DPD_DICTS):
    """rows as python lists of raw values / None; DPD cells hold RAW dictionary values"""
    dpd = 0
    cols = []
    for feat in model.features:
        t = feat["type"]
        if t == "bb":
            col = rng.integers(0, 2, n_rows).astype(object)
        elif t == "dd":
            col = rng.integers(0, len(feat["alphas"]), n_rows).astype(object)
        elif t == "dpd":
            col = np.asarray(dpd_dicts[dpd], dtype=object)[rng.integers(0, len(dpd_dicts[dpd]), n_rows)]
            dpd += 1
        elif t == "gp":
            col = rng.poisson(3.0, n_rows).astype(object)
        else:
            col = rng.normal(size=n_rows).astype(np.float32).astype(object)
        col[rng.random(n_rows) >= density] = None
        cols.append(col)
    return [[cols[f][r] for f in range(len(cols))] for r in range(n_rows)]


def fill_value(pv, model, cells, form):
    """cells: {feature: raw value}; form: 'auto' | 'dense' | 'sparse' | 'all'"""
    F = model.n_features
    feats = sorted(cells)
    if form == "auto":
        form = "none" if not feats else "all" if len(feats) == F else "sparse" if len(feats) < 0.1 * F else "dense"
    if form == "all":
        assert len(feats) == F
        pv.observed.sparsity = 3
    elif form == "dense":
        pv.observed.sparsity = 2
        pv.observed.dense.extend([f in cells for f in range(F)])
    elif form == "sparse":
        pv.observed.sparsity = 1
        pv.observed.sparse.extend(feats)
    else:
        assert not feats
        pv.observed.sparsity = 0
    for f in feats:
        t = model.features[f]["type"]
        if t == "bb":
            pv.booleans.append(bool(cells[f]))
        elif t == "nich":
            pv.reals.append(float(cells[f]))
        else:
            pv.counts.append(int(cells[f]))


def google_rows(model, table, forms, rowid0=1000):
    out = []
    for r, row in enumerate(table):
        msg = M()["Row"]()
        msg.id = rowid0 + 3 * r
        cells = {f: v for f, v in enumerate(row) if v is not None}
        fill_value(msg.diff.pos, model, cells, forms[r % len(forms)] if len(cells) < model.n_features or
                   forms[r % len(forms)] != "sparse" else "sparse")
        msg.diff.neg.observed.sparsity = 0
        out.append(msg)
    return out


def expected_csr(model, table, dpd_dicts=DPD_DICTS):
    index = []
    dpd = 0
    for feat in model.features:
        if feat["type"] == "dpd":
            index.append({v: i for i, v in enumerate(dpd_dicts[dpd])})
            dpd += 1
        else:
            index.append(None)
    ptr, feats, vals = [0], [], []
    for row in table:
        for f, v in enumerate(row):
            if v is None:
                continue
            t = model.features[f]["type"]
            raw = index[f][v] if t == "dpd" else int(np.float32(v).view(np.uint32)) if t == "nich" else int(v)
            feats.append(f)
            vals.append(raw)
        ptr.append(len(feats))
    return np.asarray(ptr, np.uint64), np.asarray(feats, np.uint32), np.asarray(vals, np.uint32)


@pytest.mark.parametrize("name", ["rows.pbs.gz", "rows.pbs"])
def test_rows_decode_every_sparsity_form(tmp_path, name):
    rng = np.random.default_rng(5)
    model = mixed_model()
    table = random_table(model, 700, rng)
    table[3] = [None] * model.n_features                       # NONE
    table[4] = random_table(model, 1, rng, density=2.0)[0]     # ALL
    table[5] = [None] * model.n_features
    table[5][9] = 2                                            # a single cell -> SPARSE under 'auto'
    msgs = google_rows(model, table, ["auto", "dense", "sparse"])
    path = tmp_path / name
    pb_schema.dump_stream(path, msgs)
    mf = pbio.ModelFile.create(model, None, DPD_DICTS)
    want_ptr, want_f, want_v = expected_csr(model, table)
    for threads in (1, 4):
        rows = pbio.rows_read(path, mf, n_threads=threads)
        assert rows.n_rows == 700 and rows.stream_rows == 700
        assert np.array_equal(rows.rowids, 1000 + 3 * np.arange(700))
        assert np.array_equal(rows.pos_ptr, want_ptr)
        assert np.array_equal(rows.pos_feature, want_f) and np.array_equal(rows.pos_value, want_v)
        assert rows.neg_ptr[-1] == 0 and not rows.row_has_tare.any()
    assert pbio.stream_stats(path) == (700, max(len(m.SerializeToString()) for m in msgs))
    # a window of the stream (InFile::set_position + a bounded read)
    win = pbio.rows_read(path, mf, first_row=100, max_rows=250)
    assert win.n_rows == 250 and win.stream_rows == 0 and win.rowids[0] == 1300
    assert np.array_equal(win.pos_ptr, want_ptr[100:351] - want_ptr[100])
    assert np.array_equal(win.pos_value, want_v[int(want_ptr[100]):int(want_ptr[350])])
    tail = pbio.rows_read(path, mf, first_row=650)
    assert tail.n_rows == 50 and tail.stream_rows == 700
    # dense columns on the host == RowBlock.from_table on dictionary indices
    cat8, wide, mask = rows.to_columns(mf)
    idx_table = []
    dmaps = [{v: i for i, v in enumerate(d)} for d in DPD_DICTS]
    dpd_feats = [f for f, ft in enumerate(model.features) if ft["type"] == "dpd"]
    for row in table:
        row = list(row)
        for j, f in enumerate(dpd_feats):
            if row[f] is not None:
                row[f] = dmaps[j][row[f]]
        idx_table.append(row)
    block = RowBlock.from_table(model, idx_table)
    assert np.array_equal(cat8, block.cat8) and np.array_equal(wide, block.wide) and np.array_equal(mask, block.mask)


def test_rows_packed_encoding_is_accepted(tmp_path):
    """a proto3-style writer packs repeated scalars; every protobuf parser accepts both forms"""
    model = Model(synth.make_features([("bb", 2), ("dd16", 1), ("nich", 1)]), [0, 0, 0, 0])
    mf = pbio.ModelFile.create(model)

    def varint(v):
        out = b""
        while v >= 0x80:
            out += bytes([v & 0x7f | 0x80])
            v >>= 7
        return out + bytes([v])

    def field(n, wt, payload):
        return varint(n << 3 | wt) + (varint(len(payload)) + payload if wt == 2 else payload)

    observed = field(1, 0, varint(2)) + field(2, 2, bytes([1, 0, 1, 1]))            # DENSE, packed bools
    value = (field(1, 2, observed) + field(2, 2, bytes([1])) + field(3, 2, varint(9)) +
             field(4, 2, struct.pack("<f", 2.5)))
    none = field(1, 2, field(1, 0, varint(0)))
    row = field(1, 0, varint(77)) + field(2, 2, field(1, 2, value) + field(2, 2, none) + field(3, 2, b""))
    path = tmp_path / "rows.pbs"
    open(path, "wb").write(struct.pack("<I", len(row)) + row)
    rows = pbio.rows_read(path, mf)
    assert rows.rowids.tolist() == [77] and rows.pos_feature.tolist() == [0, 2, 3]
    assert rows.pos_value.tolist() == [1, 9, int(np.float32(2.5).view(np.uint32))]
    g = pb_schema.load_stream(path, M()["Row"])[0]       # google agrees on what the bytes mean
    assert list(g.diff.pos.observed.dense) == [True, False, True, True] and list(g.diff.pos.counts) == [9]


def test_rows_write_is_read_by_google_and_round_trips(tmp_path):
    rng = np.random.default_rng(11)
    model = mixed_model()
    mf = pbio.ModelFile.create(model, None, DPD_DICTS)
    table = random_table(model, 300, rng, density=0.4)
    table[0] = [None] * model.n_features
    table[1] = random_table(model, 1, rng, density=2.0)[0]
    ptr, feats, vals = expected_csr(model, table)
    rows = pbio.Rows(np.arange(300) * 2, np.zeros(300, np.uint8), ptr, feats, vals, np.zeros(301, np.uint64), [])
    path = tmp_path / "rows.pbs.gz"
    pbio.rows_write(path, mf, rows)
    msgs = pb_schema.load_stream(path, M()["Row"])
    assert len(msgs) == 300 and all(m.IsInitialized() for m in msgs)
    assert msgs[0].diff.pos.observed.sparsity == 0 and msgs[1].diff.pos.observed.sparsity == 3
    for r in (1, 7, 150):
        want = [v for v in table[r] if v is not None]
        m = msgs[r].diff.pos
        got = list(m.booleans) + list(m.counts) + list(m.reals)
        assert len(got) == len(want) and np.allclose(np.asarray(got, np.float64), np.asarray(want, np.float64))
    back = pbio.rows_read(path, mf)
    assert np.array_equal(back.pos_ptr, ptr) and np.array_equal(back.pos_feature, feats)
    assert np.array_equal(back.pos_value, vals) and np.array_equal(back.rowids, rows.rowids)


def test_rows_many_batches_and_threads_agree(tmp_path):
    """more rows than one decode batch (65536), many threads == one thread"""
    rng = np.random.default_rng(2)
    model = Model(synth.make_features([("bb", 4), ("dd16", 2)]), [0] * 6)
    mf = pbio.ModelFile.create(model)
    R = 70000
    obs = rng.random((R, 6)) < 0.5
    val = np.concatenate([rng.integers(0, 2, (R, 4)), rng.integers(0, 16, (R, 2))], axis=1)
    ptr = np.zeros(R + 1, np.uint64)
    ptr[1:] = np.cumsum(obs.sum(axis=1))
    r_idx, f_idx = np.nonzero(obs)
    rows = pbio.Rows(np.arange(R), np.zeros(R, np.uint8), ptr, f_idx, val[r_idx, f_idx], np.zeros(R + 1, np.uint64), [])
    path = tmp_path / "rows.pbs.gz"
    pbio.rows_write(path, mf, rows)
    one = pbio.rows_read(path, mf, n_threads=1)
    many = pbio.rows_read(path, mf, n_threads=8)
    for a in (one, many):
        assert a.n_rows == R and np.array_equal(a.pos_ptr, ptr)
        assert np.array_equal(a.pos_feature, f_idx) and np.array_equal(a.pos_value, val[r_idx, f_idx])
        assert np.array_equal(a.rowids, np.arange(R))


def test_rows_errors(tmp_path):
    model = mixed_model()
    mf = pbio.ModelFile.create(model, None, DPD_DICTS)
    path = tmp_path / "rows.pbs"

    def one_row(mutate):
        msg = M()["Row"]()
        msg.id = 1
        msg.diff.pos.observed.sparsity = 0
        msg.diff.neg.observed.sparsity = 0
        mutate(msg)
        pb_schema.dump_stream(path, [msg])

    def dd_out_of_range(m):
        fill_value(m.diff.pos, model, {3: 16}, "sparse")          # feature 3 is a dd16
    one_row(dd_out_of_range)
    with pytest.raises(_abi.LoomError, match="out of range"):
        pbio.rows_read(path, mf)

    def dpd_unknown(m):
        fill_value(m.diff.pos, model, {7: 8}, "sparse")           # 8 is not in DPD_DICTS[0]
    one_row(dpd_unknown)
    with pytest.raises(_abi.LoomError, match="dictionary"):
        pbio.rows_read(path, mf)

    def wrong_dense(m):
        m.diff.pos.observed.sparsity = 2
        m.diff.pos.observed.dense.extend([True, False])
    one_row(wrong_dense)
    with pytest.raises(_abi.LoomError, match="wrong size"):
        pbio.rows_read(path, mf)

    def too_many_values(m):
        fill_value(m.diff.pos, model, {0: 1}, "sparse")
        m.diff.pos.booleans.append(True)
    one_row(too_many_values)
    with pytest.raises(_abi.LoomError, match="more packed values"):
        pbio.rows_read(path, mf)

    def second_tare(m):
        m.diff.tares.append(1)
    one_row(second_tare)
    with pytest.raises(_abi.LoomError, match="tare"):
        pbio.rows_read(path, mf)

    one_row(lambda m: None)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw + raw[:len(raw) - 2])              # second frame cut short
    with pytest.raises(_abi.LoomError, match="truncated"):
        pbio.rows_read(path, mf)
    with pytest.raises(_abi.LoomError, match="failed to open"):
        pbio.rows_read(tmp_path / "nope.pbs.gz", mf)


def test_tared_rows_reconstruct_the_table(tmp_path):
    """sparsified rows (Diff{pos, neg, tares:[0]}) + tares file -> the dense block they came from;
    the worked example of doc/adapting.md:323-414 in file form"""
    rng = np.random.default_rng(8)
    feats = synth.make_features([("bb", 6), ("dd16", 3), ("gp", 2), ("nich", 2)])
    model = Model(feats, [0] * len(feats))
    mf = pbio.ModelFile.create(model)
    R = 400
    table = random_table(model, R, rng, density=0.9)
    for row in table:                         # make some columns mostly constant so they get a tare
        for f in (0, 1, 2, 6, 9):
            if row[f] is not None and rng.random() < 0.9:
                row[f] = 1 if f < 6 else 3
    block = RowBlock.from_table(model, table)
    tare_obs, tare_val = tare.make_tare(model, block)
    assert tare_obs.sum() >= 4
    d = tare.sparsify(model, block, tare_obs, tare_val)
    # the files a loom front end would hand over
    tares_path, rows_path = tmp_path / "tares.pbs.gz", tmp_path / "rows.pbs.gz"
    tmsg = M()["ProductValue"]()
    fill_value(tmsg, model, {f: (np.uint32(tare_val[f]).view(np.float32) if feats[f]["type"] == "nich" else tare_val[f])
                             for f in range(len(feats)) if tare_obs[f]}, "auto")
    pb_schema.dump_stream(tares_path, [tmsg])
    msgs = []
    for r in range(R):
        msg = M()["Row"]()
        msg.id = r
        p0, p1 = int(d["pos_ptr"][r]), int(d["pos_ptr"][r + 1])
        cells = {}
        for f, v in zip(d["pos_feature"][p0:p1], d["pos_value"][p0:p1]):
            cells[int(f)] = np.uint32(v).view(np.float32) if feats[f]["type"] == "nich" else int(v)
        fill_value(msg.diff.pos, model, cells, "auto")
        n0, n1 = int(d["neg_ptr"][r]), int(d["neg_ptr"][r + 1])
        fill_value(msg.diff.neg, model, {int(f): int(tare_val[f]) for f in d["neg_feature"][n0:n1]}, "auto")
        msg.diff.tares.append(0)
        msgs.append(msg)
    pb_schema.dump_stream(rows_path, msgs)
    to, tv, n = pbio.tares_read(tares_path, mf)
    assert n == 1 and np.array_equal(to, tare_obs) and np.array_equal(tv[to == 1], tare_val[tare_obs == 1])
    rows = pbio.rows_read(rows_path, mf)
    assert rows.row_has_tare.all()
    assert np.array_equal(rows.pos_ptr, d["pos_ptr"]) and np.array_equal(rows.pos_feature, d["pos_feature"])
    assert np.array_equal(rows.pos_value, d["pos_value"])
    assert np.array_equal(rows.neg_ptr, d["neg_ptr"]) and np.array_equal(rows.neg_feature, d["neg_feature"])
    cat8, wide, mask = rows.to_columns(mf, to, tv)
    assert np.array_equal(cat8, block.cat8) and np.array_equal(wide, block.wide) and np.array_equal(mask, block.mask)
    # tares_write -> google
    out = tmp_path / "tares.out.pbs"
    pbio.tares_write(out, mf, to, tv)
    (g,) = pb_schema.load_stream(out, M()["ProductValue"])
    assert g == tmsg
    with pytest.raises(_abi.LoomError, match="no tares"):
        rows.to_columns(mf)


# ------------------------------------------------------------------------------------------ assignments / groups / log
def test_assignments_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    rowids = rng.integers(0, 2 ** 50, 500).astype(np.uint64)
    gids = rng.integers(0, 300, (500, 4)).astype(np.uint32)
    path = tmp_path / "assign.pbs.gz"
    pbio.assign_write(path, rowids, gids)
    msgs = pb_schema.load_stream(path, M()["Assignment"])
    assert [m.rowid for m in msgs] == rowids.tolist() and list(msgs[17].groupids) == gids[17].tolist()
    # a stream written by google
    pb_schema.dump_stream(path, msgs[:100])
    r2, g2 = pbio.assign_read(path)
    assert np.array_equal(r2, rowids[:100]) and np.array_equal(g2, gids[:100])
    msgs[5].groupids.append(1)
    pb_schema.dump_stream(path, msgs[:10])
    with pytest.raises(_abi.LoomError, match="kind count"):
        pbio.assign_read(path)
    pb_schema.dump_stream(path, [])
    r3, g3 = pbio.assign_read(path)
    assert len(r3) == 0


def google_groups(model, mf, fids, counts, ints, flts, order):
    dicts = mf.dpd_values()
    dpd_of = {f: i for i, f in enumerate(f for f, ft in enumerate(model.features) if ft["type"] == "dpd")}
    out = []
    for g in order:
        msg = M()["ProductModel.Group"]()
        msg.count = int(counts[g])
        io = fo = 0
        for f in fids:
            feat = model.features[f]
            t = feat["type"]
            if t == "bb":
                s = msg.bb.add()
                s.heads, s.tails = int(ints[io, g]), int(ints[io + 1, g])
            elif t == "dd":
                msg.dd.add().counts.extend(int(x) for x in ints[io:io + len(feat["alphas"]), g])
            elif t == "dpd":
                s = msg.dpd.add()
                for v in range(len(feat["betas"])):
                    if ints[io + v, g]:
                        s.keys.append(int(dicts[dpd_of[f]][v]))
                        s.values.append(int(ints[io + v, g]))
            elif t == "gp":
                s = msg.gp.add()
                s.count, s.sum, s.log_prod = int(ints[io, g]), int(ints[io + 1, g]), float(flts[fo, g])
            else:
                s = msg.nich.add()
                s.count, s.mean, s.count_times_variance = int(ints[io, g]), float(flts[fo, g]), float(flts[fo + 1, g])
            io += Model.int_rows(feat)
            fo += Model.flt_rows(feat)
        out.append(msg)
    return out


def test_groups_round_trip(tmp_path):
    rng = np.random.default_rng(4)
    model = mixed_model()
    mf = pbio.ModelFile.create(model, None, DPD_DICTS)
    fids = model.kind_features(0) + model.kind_features(1)
    fids.sort()
    G = 9
    n_int = sum(Model.int_rows(model.features[f]) for f in fids)
    n_flt = sum(Model.flt_rows(model.features[f]) for f in fids)
    counts = rng.integers(1, 50, G).astype(np.uint32)
    ints = rng.integers(0, 20, (n_int, G)).astype(np.uint32)
    flts = rng.normal(size=(n_flt, G)).astype(np.float32).astype(np.float64)
    io = 0
    for f in fids:                       # make the derived count_sum rows consistent
        feat = model.features[f]
        if feat["type"] in ("dd", "dpd"):
            dim = Model.int_rows(feat) - 1
            ints[io + dim] = ints[io:io + dim].sum(axis=0)
        io += Model.int_rows(feat)
    order = np.argsort(-counts.astype(np.int64), kind="stable").astype(np.uint32)[:7]   # descending count, 2 dropped
    path = tmp_path / pbio.mixture_path("", 3).lstrip("/")
    pbio.groups_write(path, mf, fids, counts, ints, flts, order)
    got = pb_schema.load_stream(path, M()["ProductModel.Group"])
    want = google_groups(model, mf, fids, counts, ints, flts, order)
    assert len(got) == 7 and all(g.IsInitialized() for g in got)
    assert got == want
    c2, i2, f2 = pbio.groups_read(path, mf, fids)
    assert np.array_equal(c2, counts[order]) and np.array_equal(i2, ints[:, order]) and np.array_equal(f2, flts[:, order])
    # wrong feature list -> loud
    with pytest.raises(_abi.LoomError, match="feature entries|counts"):
        pbio.groups_read(path, mf, fids[:-1])


def test_log_stream(tmp_path):
    path = tmp_path / "log.pbs.gz"
    log = pbio.C.c_void_p()
    pbio._check(pbio.lib().lbio_log_open(os.fsencode(path), 0, pbio.C.byref(log)))
    kind_py = np.asarray([[2.0, 0.1], [3.0, 0.2]], np.float32)
    fc, cc = np.asarray([3, 5], np.uint32), np.asarray([10, 20], np.uint32)
    for it in range(3):
        e = pbio.LogEntry()
        e.iter, e.n_kinds = it, 2
        e.topology_py[0], e.topology_py[1] = 1.5, 0.25
        e.kind_py, e.feature_counts, e.category_counts = (_abi.ptr(kind_py, _abi._f32p), _abi.ptr(fc, _abi._u32p),
                                                          _abi.ptr(cc, _abi._u32p))
        e.score, e.kl_divergence, e.assigned_object_count = -100.5 - it, 0.25, 40 + it
        e.kind_total_count, e.kind_change_count = 8, 2
        pbio._check(pbio.lib().lbio_log_write(log, pbio.C.byref(e)))
    pbio.lib().lbio_log_close(log)
    msgs = pb_schema.load_stream(path, M()["LogMessage"])
    assert len(msgs) == 3 and all(m.IsInitialized() for m in msgs)
    last = msgs[2].args
    assert last.iter == 2 and last.scores.score == -102.5 and last.scores.assigned_object_count == 42
    assert list(last.summary.feature_counts) == [3, 5] and list(last.summary.category_counts) == [10, 20]
    assert last.summary.model_hypers.alpha == 1.5 and last.summary.kind_hypers[1].d == np.float32(0.2)
    assert last.kernel_status.kind.change_count == 2 and msgs[0].timestamp_usec > 0
