"""loom_b200.query -- the loom.query front end (loom/query.py) -- against the server process built on the oracle: the
checks of loom/test/test_query_math.py (score of the empty row, samples against scores, entropy against a Monte-Carlo
estimate from the client's own samples and scores, mutual information) and loom/test/test_query.py (batch_score,
score_derivative) read through the client API."""
import numpy as np
import pytest

from loom_b200 import query
from test_pbio import io_lib  # noqa: F401
from test_query import MIN_P, chi_square_p
from test_query_server import cells_of, query_on_oracle, some_rows, store  # noqa: F401
from test_z_loom_infer import infer_on_oracle  # noqa: F401


def typed_row(model, row):
    """a data row as loom.query takes it: bool / int / float / None by the feature's class"""
    out = []
    for f, v in enumerate(row):
        t = model.features[f]["type"]
        out.append(None if v is None else bool(v) if t == "bb" else float(v) if t == "nich" else int(v))
    return out


@pytest.fixture()
def server(store, query_on_oracle):  # noqa: F811
    root, model, block, rowids = store
    with query.get_server(root, config={"seed": 5}, exe=query_on_oracle) as s:
        yield s, model, block, rowids


def test_score_none_is_zero_and_batch_score_matches_score(server):
    s, model, block, rowids = server
    F = model.n_features
    assert abs(s.score([None] * F)) < 1e-3                                   # test_query_math.py:49-54
    rows = [typed_row(model, r) for r in some_rows(model, block, range(25))]
    one_by_one = [s.score(r) for r in rows[:5]]
    batched = list(s.batch_score(rows))                                      # more rows than the send-ahead buffer
    assert len(batched) == 25 and batched[:5] == one_by_one
    assert all(np.isfinite(batched))


def test_samples_match_scores(server):
    """test_query_math.py:65-92 through the client: frequencies of sampled values against exp(score differences)"""
    s, model, block, rowids = server
    F = model.n_features
    cond = typed_row(model, some_rows(model, block, [2])[0])
    bb = [f for f in range(F) if model.features[f]["type"] == "bb"][:2]
    for f in bb:
        cond[f] = None
    to_sample = [f in bb for f in range(F)]
    n = 2000
    samples = s.sample(to_sample, cond, n)
    assert len(samples) == n
    for smp in samples[:20]:                                                 # conditioning values are carried over
        assert all(smp[f] == cond[f] for f in range(F) if f not in bb)
        assert all(isinstance(smp[f], bool) for f in bb)
    base = s.score(cond)
    combos = [(a, b) for a in (False, True) for b in (False, True)]
    probs = []
    for a, b in combos:
        row = list(cond)
        row[bb[0]], row[bb[1]] = a, b
        probs.append(np.exp(s.score(row) - base))
    assert abs(sum(probs) - 1) < 1e-4
    counts = [sum(1 for smp in samples if (smp[bb[0]], smp[bb[1]]) == c) for c in combos]
    assert chi_square_p(counts, probs) > MIN_P


def test_entropy_and_mutual_information(server):
    """test_query_math.py:111-142: the server's entropy against an estimate from the client's own samples and scores,
    within 4 sigma; mutual information is non-negative within its error and adds up from the entropies"""
    s, model, block, rowids = server
    F = model.n_features
    bb = [f for f in range(F) if model.features[f]["type"] == "bb"]
    fs1, fs2 = frozenset([bb[0]]), frozenset([bb[1], bb[2]])
    n = 1500
    ent = s.entropy([fs1], [fs2], None, n)
    assert set(ent) == {frozenset(), fs1, fs2, fs1 | fs2}
    assert abs(ent[frozenset()].mean) < 1e-5
    for fs in (fs1, fs2, fs1 | fs2):
        to_sample = [f in fs for f in range(F)]
        draws = s.sample(to_sample, None, 300)
        logp = np.array(list(s.batch_score([[v if f in fs else None for f, v in enumerate(d)] for d in draws])))
        mc_mean, mc_var = -logp.mean(), logp.var() / len(logp)
        assert abs(ent[fs].mean - mc_mean) < 4 * np.sqrt(ent[fs].variance + mc_var) + 1e-6, fs
        assert 0 < ent[fs].mean <= len(fs) * np.log(2) + 1e-3                # booleans: at most one bit each
    mi = s.mutual_information(fs1, fs2, sample_count=n)
    mi2 = s.mutual_information(fs1, fs2, entropys=ent)
    assert mi2.mean == pytest.approx(ent[fs1].mean + ent[fs2].mean - ent[fs1 | fs2].mean)
    assert mi.mean > -4 * np.sqrt(mi.variance) and abs(mi.mean - mi2.mean) < 4 * np.sqrt(mi.variance + mi2.variance)


def test_score_derivative(server):
    """loom/test/test_query.py:181-204"""
    s, model, block, rowids = server
    rows = [typed_row(model, r) for r in some_rows(model, block, range(3))]
    results = s.score_derivative(rows[0], rows[:2])
    assert len(results) == 2 and {i for i, _ in results} == {0, 1}
    results = s.score_derivative(rows[0], score_rows=None)
    assert len(results) == block.n_rows and sorted(i for i, _ in results) == sorted(rowids.tolist())
    assert len(s.score_derivative(rows[0], score_rows=None, row_limit=1)) == 1


def test_errors_raise(server):
    s, model, block, rowids = server
    F = model.n_features
    with pytest.raises(Exception, match="invalid request.score.data"):
        s.score([True] * (F + 3))                                            # a mask of the wrong size
    assert np.isfinite(s.score([None] * F))                                  # the server carries on
    with pytest.raises(Exception, match="invalid request.entropy.sample_count"):
        s.entropy([[0]], [[1]], None, 1)


def test_missing_binary_is_loud(tmp_path):
    with pytest.raises(RuntimeError, match="no Python fallback"):
        query.get_server(tmp_path, exe=str(tmp_path / "nope"))
