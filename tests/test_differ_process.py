"""loom_tare and loom_sparsify -- the processes loom.runner starts before inference (src/tare.cc, src/sparsify.cc) --
built against the oracle's lo_ ABI and driven through files: a schema row, a dense rows stream written by
google.protobuf; outputs parsed with google.protobuf and compared with the numpy restatement of loom::Differ
(loom_b200/tare.py).  The product binaries run on the device in tests/test_zz_query_gpu.py."""
import subprocess

import numpy as np
import pytest

import pb_schema
from loom_b200 import RowBlock, pack_mask, pbio, synth, tare
from test_pbio import M, io_lib  # noqa: F401  (io_lib: autouse fixture builds the codec)
from test_query_server import build_on_oracle, value_cells


@pytest.fixture(scope="module")
def tare_on_oracle(tmp_path_factory, oracle_abi):
    return build_on_oracle(tmp_path_factory, "loom_tare.cc", "loom_tare_oracle")


@pytest.fixture(scope="module")
def sparsify_on_oracle(tmp_path_factory, oracle_abi):
    return build_on_oracle(tmp_path_factory, "loom_sparsify.cc", "loom_sparsify_oracle")


def dataset(n_rows=500, seed=2, dominant=True):
    rng = np.random.default_rng(seed)
    model, rows, _ = synth.generate(n_rows, [("bb", 5), ("dd16", 2), ("gp", 3), ("nich", 2)], density=0.8, seed=seed)
    if dominant:
        rows.cat8[:3] = rng.random((3, n_rows)) < 0.05                 # sparse booleans
        rows.cat8[5] = np.where(rng.random(n_rows) < 0.8, 4, rows.cat8[5])
        rows.wide[0] = np.where(rng.random(n_rows) < 0.75, 0, rows.wide[0])
        obs = rows.observed()
        obs[:3] = True
        obs[5] |= rng.random(n_rows) < 0.9
        obs[7] |= rng.random(n_rows) < 0.9
        rows = RowBlock(n_rows, rows.cat8, rows.wide, pack_mask(obs))
    else:
        rows.cat8[:5] = rng.random((5, n_rows)) < 0.5
        rows.cat8[5:7] = rng.integers(0, 16, (2, n_rows))
        rows.wide[:3] = rng.integers(0, 9, (3, n_rows))
    return model, rows


def write_files(d, model, rows, seed=0):
    schema_row = M()["ProductValue"]()                                  # loom/format.py:79-101
    schema_row.observed.sparsity = 2
    for f in model.features:
        schema_row.observed.dense.append(True)
        if f["type"] == "bb":
            schema_row.booleans.append(False)
        elif f["type"] == "nich":
            schema_row.reals.append(0.0)
        else:
            schema_row.counts.append(0)
    pb_schema.dump_message(d / "schema.pb.gz", schema_row)
    mf = pbio.ModelFile.create(model)
    rowids = np.random.default_rng(seed).permutation(10 * rows.n_rows)[:rows.n_rows].astype(np.uint64)
    pbio.rows_write(d / "rows.pbs.gz", mf, pbio.Rows.from_block(rows, rowids))
    return mf, rowids


def run(cmd, window_rows=None):
    import os
    env = dict(os.environ)
    if window_rows:
        env["LOOM_B200_WINDOW_ROWS"] = str(window_rows)
    proc = subprocess.run([str(c) for c in cmd], capture_output=True, timeout=600, env=env)
    return proc.returncode, proc.stderr.decode()


def passes_case(tare_exe, sparsify_exe, d, window_rows=None):
    model, rows = dataset()
    mf, rowids = write_files(d, model, rows)
    assert run([tare_exe, d / "schema.pb.gz", d / "rows.pbs.gz", d / "tares.pbs.gz"], window_rows) == (0, "")
    want_obs, want_val = tare.make_tare(model, rows)
    assert want_obs.sum() >= 5
    tares = pb_schema.load_stream(d / "tares.pbs.gz", M()["ProductValue"])
    assert len(tares) == 1 and tares[0].IsInitialized()
    cells = value_cells(model, tares[0])
    assert sorted(cells) == list(np.flatnonzero(want_obs)) and all(cells[f] == want_val[f] for f in cells)
    assert run([sparsify_exe, d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "diffs.pbs.gz"], window_rows) == (0, "")
    want = tare.sparsify(model, rows, want_obs, want_val)
    got = pbio.rows_read(d / "diffs.pbs.gz", mf)
    assert np.array_equal(got.rowids, rowids) and got.row_has_tare.all()
    for key in ("pos_ptr", "pos_feature", "pos_value", "neg_ptr", "neg_feature"):
        assert np.array_equal(getattr(got, key), want[key]), key
    # on the wire: neg holds the tare's values (src/differ.cc:277-286), tares = [0], both parts validly packed
    msgs = pb_schema.load_stream(d / "diffs.pbs.gz", M()["Row"])
    for r in (0, 1, 2, len(msgs) - 1):
        assert list(msgs[r].diff.tares) == [0]
        neg = value_cells(model, msgs[r].diff.neg)
        assert all(v == want_val[f] for f, v in neg.items())
        value_cells(model, msgs[r].diff.pos)
    return model, rows, mf


def test_tare_then_sparsify(tare_on_oracle, sparsify_on_oracle, tmp_path):
    passes_case(tare_on_oracle, sparsify_on_oracle, tmp_path)


@pytest.mark.parametrize("window_rows", [128, 250, 500])
def test_streams_longer_than_one_block(tare_on_oracle, sparsify_on_oracle, tmp_path, window_rows):
    """windows of the stream: summaries add up across blocks, the threshold is over all rows, the diff stream is
    the concatenation (500 rows: ragged last window, exact multiple, one full window)"""
    passes_case(tare_on_oracle, sparsify_on_oracle, tmp_path, window_rows)


def test_no_dominant_value_means_no_tare(tare_on_oracle, sparsify_on_oracle, tmp_path):
    d = tmp_path
    model, rows = dataset(dominant=False)
    mf, rowids = write_files(d, model, rows)
    assert run([tare_on_oracle, d / "schema.pb.gz", d / "rows.pbs.gz", d / "tares.pbs.gz"]) == (0, "")
    assert pb_schema.load_stream(d / "tares.pbs.gz", M()["ProductValue"]) == []        # src/tare.cc:66-69
    assert run([sparsify_on_oracle, d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "diffs.pbs.gz"]) == (0, "")
    got, src = pbio.rows_read(d / "diffs.pbs.gz", mf), pbio.rows_read(d / "rows.pbs.gz", mf)
    assert not got.row_has_tare.any() and int(got.neg_ptr[-1]) == 0
    for key in ("rowids", "pos_ptr", "pos_feature", "pos_value"):
        assert np.array_equal(getattr(got, key), getattr(src, key)), key


def test_errors(tare_on_oracle, sparsify_on_oracle, tmp_path):
    d = tmp_path
    model, rows, mf = passes_case(tare_on_oracle, sparsify_on_oracle, d)
    rc, err = run([tare_on_oracle, d / "schema.pb.gz", d / "diffs.pbs.gz", d / "tares2.pbs.gz"])
    assert rc != 0 and "row is already sparsified" in err                              # src/differ.cc:98
    rc, err = run([sparsify_on_oracle, d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "rows.pbs.gz"])
    assert rc != 0 and "in-place sparsify is not supported" in err                     # src/differ.cc:156-160
    rc, err = run([tare_on_oracle, d / "missing.pb.gz", d / "rows.pbs.gz", d / "t.pbs.gz"])
    assert rc != 0 and "failed to open input file" in err
    rc, err = run([tare_on_oracle, "too", "few"])
    assert rc == 1 and "Usage: loom_tare SCHEMA_ROW_IN ROWS_IN TARES_OUT" in err


# ------------------------------------------------------------------------------------------ loom_shuffle
def test_shuffle_is_a_seeded_permutation_and_chunk_size_invariant(tmp_path):
    """loom/test/test_shuffle.py:41-82 through the product binary (host I/O only, no device): the output is a
    permutation of the input's messages, the same for any memory target, different for another seed"""
    import os
    from test_z_loom_infer import HOST
    exe = os.path.join(HOST, "loom_shuffle")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST, "loom_shuffle"], stdout=subprocess.DEVNULL)
    model, rows = dataset(n_rows=300)
    mf, rowids = write_files(tmp_path, model, rows)
    outs = {}
    for tag, seed, mem in (("a", 7, 4e9), ("b", 7, 5000), ("c", 7, 1), ("d", 8, 4e9)):
        out = tmp_path / (tag + ".pbs.gz")
        assert run([exe, tmp_path / "rows.pbs.gz", out, seed, mem]) == (0, "")
        outs[tag] = pbio.rows_read(out, mf).rowids.tolist()
    assert sorted(outs["a"]) == sorted(rowids.tolist()) and outs["a"] != rowids.tolist()
    assert outs["a"] == outs["b"] == outs["c"]                    # one message per chunk at the smallest target
    assert outs["d"] != outs["a"] and sorted(outs["d"]) == sorted(outs["a"])
    rc, err = run([exe, tmp_path / "rows.pbs.gz", tmp_path / "rows.pbs.gz"])
    assert rc != 0 and "cannot shuffle file in-place" in err       # src/shuffle.hpp:46-48
    rc, err = run([exe, "-", tmp_path / "x.pbs"])
    assert rc != 0 and "shuffle input is not a file" in err
    out = tmp_path / "default.pbs.gz"                              # SEED and TARGET_MEM_BYTES are optional
    assert run([exe, tmp_path / "rows.pbs.gz", out]) == (0, "") and pbio.stream_stats(out)[0] == 300
    pbio.shuffle_stream(tmp_path / "rows.pbs.gz", tmp_path / "py.pbs.gz", seed=7)
    assert pbio.rows_read(tmp_path / "py.pbs.gz", mf).rowids.tolist() == outs["a"]


def test_tasks_mirror_equals_the_processes(tare_on_oracle, sparsify_on_oracle, oracle_abi, tmp_path):
    """loom_b200.tasks.tare / sparsify / shuffle (loom.runner's functions, in process) write the files the processes write"""
    from loom_b200 import tasks
    d = tmp_path
    model, rows, mf = passes_case(tare_on_oracle, sparsify_on_oracle, d)
    tasks.tare(d / "schema.pb.gz", d / "rows.pbs.gz", d / "py.tares.pbs.gz", abi=oracle_abi)
    tasks.sparsify(d / "schema.pb.gz", d / "py.tares.pbs.gz", d / "rows.pbs.gz", d / "py.diffs.pbs.gz", abi=oracle_abi)
    import gzip
    for name in ("tares.pbs.gz", "diffs.pbs.gz"):
        assert gzip.open(d / ("py." + name)).read() == gzip.open(d / name).read(), name     # byte for byte
    tasks.shuffle(d / "diffs.pbs.gz", d / "py.shuffled.pbs.gz", seed=3)
    assert sorted(pbio.rows_read(d / "py.shuffled.pbs.gz", mf).rowids.tolist()) == sorted(pbio.rows_read(d / "diffs.pbs.gz", mf).rowids.tolist())
    with pytest.raises(Exception, match="already sparsified"):
        tasks.tare(d / "schema.pb.gz", d / "diffs.pbs.gz", d / "x.pbs.gz", abi=oracle_abi)
    with pytest.raises(Exception, match="in-place"):
        tasks.sparsify(d / "schema.pb.gz", d / "tares.pbs.gz", d / "rows.pbs.gz", d / "rows.pbs.gz", abi=oracle_abi)
