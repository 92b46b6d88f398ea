"""Pins of the oracle against pieces of the REFERENCE ITSELF compiled in the build container (oracle/Makefile `ref`,
oracle/_ref/): the committed fixtures tests/golden/reference_*.json hold what the reference's own code printed, the
generators sit beside them.  (The schedule classes are pinned the same way in tests/test_infer_driver.py.)"""
import ctypes
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _oracle():
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
    lib.oracle_block_py.restype = None
    lib.oracle_block_py.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]
    return lib


def _likelihoods(scores):
    """scores_to_likelihoods as lo_kind_run applies it (oracle_structure.c; src/kind_proposer.cc:347), float64"""
    s = np.asarray(scores, np.float32).astype(np.float64)
    return np.ascontiguousarray(np.exp(s - s.max(axis=1, keepdims=True)))


def test_block_pitman_yor_reproduces_the_reference_draw_for_draw():
    """src/kind_proposer.cc compiled unmodified (KindProposer::infer_assignments -> BlockPitmanYorSampler::run) against
    the oracle's oracle_block_py on the same scores, start and uniforms: every draw of every iteration must be the
    same kind.  The reference computes in float32, the oracle in float64: a uniform that lands within float rounding
    of a boundary of the CDF could legitimately differ, so a case may diverge only where the two cumulative sums are
    that close -- none of the 12 200 draws of the fixture does, and the test demands exact equality."""
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_kind_proposer.json")))
    lib = _oracle()
    assert len(fixture["cases"]) >= 8
    draws = 0
    for case in fixture["cases"]:
        F, K, iterations = case["F"], case["K"], case["iterations"]
        lik = _likelihoods(case["scores"])
        # the oracle returns the end state of a run; the uniform of a draw depends only on (seed, iteration, feature),
        # so a run of `it` iterations is a prefix of the full run and yields the state after every iteration
        assign = np.array(case["initial"], np.uint32)
        lib.oracle_block_py(case["alpha"], case["d"], F, K, lik.ctypes.data, assign.ctypes.data, iterations, case["seed"])
        assert assign.tolist() == case["assignments"], (F, K, case["seed"])
        picks = np.array(case["picks"], np.uint32).reshape(iterations, F)
        assert picks[-1].tolist() == case["assignments"]
        for it in range(1, iterations):                       # the state after every iteration, not only the last
            a = np.array(case["initial"], np.uint32)
            lib.oracle_block_py(case["alpha"], case["d"], F, K, lik.ctypes.data, a.ctypes.data, it, case["seed"])
            assert a.tolist() == picks[it - 1].tolist(), (F, K, case["seed"], it)
        draws += picks.size
    assert draws >= 10000


def test_block_pitman_yor_totals_match_the_reference():
    """The normaliser the reference draws against (sum_k prior_k * likelihood_k, float32) recomputed in float64 from
    the fixture's own pick sequence with the oracle's formulas (prior = n_k - d, or (alpha + d * nonempty) / empty
    for an empty kind: src/kind_proposer.cc:174-226): within float32 rounding."""
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_kind_proposer.json")))
    worst = 0.0
    for case in fixture["cases"]:
        F, K, iterations, alpha, d = case["F"], case["K"], case["iterations"], case["alpha"], case["d"]
        lik = _likelihoods(case["scores"])
        assign = list(case["initial"])
        counts = np.bincount(assign, minlength=K).astype(np.int64)
        totals = np.array(case["totals"]).reshape(iterations, F)
        picks = np.array(case["picks"]).reshape(iterations, F)
        for it in range(iterations):
            for f in range(F):
                counts[assign[f]] -= 1
                empty = int((counts == 0).sum())
                prior = np.where(counts > 0, counts - d, (alpha + d * (K - empty)) / max(empty, 1))
                total = float((prior * lik[f]).sum())
                worst = max(worst, abs(total - totals[it, f]) / total)
                assign[f] = int(picks[it, f])
                counts[assign[f]] += 1
    assert worst < 2e-6, worst


# ------------------------------------------------------------------------------------------ loom::Differ (tare, sparsify)
def _differ_cases():
    return json.load(open(os.path.join(HERE, "golden", "reference_differ.json")))["cases"]


def _model_of(types):
    from loom_b200 import synth
    from loom_b200.engine import Model
    feats = synth.make_features([tuple(t) for t in types])
    return Model(feats, np.zeros(len(feats), np.uint32), kind_py=[synth.CLUSTERING], topology=synth.CLUSTERING)


def _cells(model, value):
    """a fixture ProductValue -> (feature ids, raw uint32 cells in the engine's encoding: NICH as float bits)"""
    feats = list(value["features"])
    vals = list(value["booleans"]) + list(value["counts"]) + [int(np.float32(x).view(np.uint32)) for x in value["reals"]]
    assert len(vals) == len(feats)
    return feats, vals


def differ_matches_the_reference(abi, tmp_path=None, cases=None, min_rows=470):
    """Differ::add_rows -> get_tare and Differ::compress_rows of the reference (tests/golden/reference_differ.json, from
    src/differ.cc compiled unmodified) against the engine's differ_* entry points on the same tables: the tare row,
    and for every row the features and values of pos and the features of neg, bit for bit.  With tmp_path the rows are
    also written through the product's codec and read back with google.protobuf: the representation of every
    pos / neg (NONE / SPARSE / DENSE / ALL, ValueSchema::normalize_small), neg's values (the tare's) and the `tares`
    field must be the reference's too."""
    from loom_b200 import Engine
    from loom_b200.engine import RowBlock
    checked = 0
    for case in (_differ_cases() if cases is None else cases):
        model = _model_of(case["types"])
        table = case["table"]
        e = Engine(abi)
        e.set_model(model)
        e.set_rows(RowBlock.from_table(model, table))
        obs, val = e.make_tare()
        want_f, want_v = _cells(model, case["tare"])
        assert np.flatnonzero(obs).tolist() == want_f, case["name"]
        assert val[want_f].tolist() == want_v, case["name"]
        d = e.sparsify(obs, val)
        for r, row in enumerate(case["rows"]):
            pf, pv = _cells(model, row["pos"])
            nf, nv = _cells(model, row["neg"])
            p0, p1 = int(d["pos_ptr"][r]), int(d["pos_ptr"][r + 1])
            n0, n1 = int(d["neg_ptr"][r]), int(d["neg_ptr"][r + 1])
            assert d["pos_feature"][p0:p1].tolist() == pf and d["pos_value"][p0:p1].tolist() == pv, (case["name"], r)
            assert d["neg_feature"][n0:n1].tolist() == nf, (case["name"], r)
            assert val[nf].tolist() == nv, (case["name"], r)        # neg carries the tare's values (differ.cc:277-286)
            assert row["tares"] == (1 if want_f else 0)
            checked += 1
        if tmp_path is not None:
            import pb_schema
            from loom_b200 import pbio
            mf = pbio.ModelFile.create(model, None, None)
            tared = np.full(len(table), 1 if want_f else 0, np.uint8)
            out = pbio.Rows(np.arange(len(table)), tared, d["pos_ptr"], d["pos_feature"], d["pos_value"], d["neg_ptr"], d["neg_feature"])
            path = tmp_path / (case["name"] + ".pbs.gz")
            pbio.rows_write(path, mf, out, tare_values=val)
            msgs = pb_schema.load_stream(path, pb_schema.messages()["Row"])
            names = {0: "NONE", 1: "SPARSE", 2: "DENSE", 3: "ALL"}
            for r, (m, row) in enumerate(zip(msgs, case["rows"])):
                assert m.id == row["id"]
                for side, got in (("pos", m.diff.pos), ("neg", m.diff.neg)):
                    want = row[side]
                    assert names[got.observed.sparsity] == want["sparsity"], (case["name"], r, side)
                    if want["sparsity"] == "SPARSE":
                        assert list(got.observed.sparse) == want["features"]
                    if want["sparsity"] == "DENSE":
                        assert np.flatnonzero(list(got.observed.dense)).tolist() == want["features"]
                    assert [int(b) for b in got.booleans] == want["booleans"] and list(got.counts) == want["counts"]
                    assert np.array_equal(np.asarray(list(got.reals), np.float32), np.asarray(want["reals"], np.float32))
                assert len(m.diff.tares) == row["tares"]
    assert checked >= min_rows


def test_oracle_differ_reproduces_the_reference(oracle_abi, tmp_path):
    differ_matches_the_reference(oracle_abi, tmp_path)


@pytest.mark.gpu
def test_cuda_differ_reproduces_the_reference(cuda_abi, tmp_path):
    differ_matches_the_reference(cuda_abi, tmp_path)


# ------------------------------------------------------------------------------------------ KindKernel bookkeeping
def test_kind_kernel_bookkeeping_reproduces_the_reference(oracle_abi):
    """src/kind_kernel.cc + src/kind_proposer.cc compiled unmodified (a KindKernel constructed over a CrossCat, try_run()
    on supplied score matrices, destroyed: tests/golden/reference_kind_kernel.json) against the oracle's
    init_featureless_kinds / kind_run on the same scores and uniforms: after EVERY step the feature -> kind map and the
    number of kinds must be the reference's -- which features moved, which kinds died, how swap-removal renumbered the
    survivors (src/kind_kernel.cc:196-206), where the fresh ephemeral kinds went.  19 try_run calls over 5 cases with
    births and deaths in most of them.  (The product's lb_kind_run / lb_init_featureless_kinds are tied to the oracle's
    by the -m gpu kind tests.)"""
    import kind_pin_util
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_kind_kernel.json")))
    runs = births = deaths = 0
    for case in fixture["cases"]:
        states = kind_pin_util.life_cycle(oracle_abi, case, lambda r, n_kinds: case["scores"][r])
        assert len(states) == len(case["steps"])
        for (step, f2k, n_kinds), want in zip(states, case["steps"]):
            assert step == want["step"]
            assert n_kinds == want["n_kinds"], (case["seed"], step)
            assert f2k == want["f2k"], (case["seed"], step)
            # the reference's own view of the same state: every kind's feature count
            assert np.bincount(f2k, minlength=n_kinds).tolist() == want["n_features"]
            if step == "try_run":
                runs += 1
                births += want["birth_count"]
                deaths += want["death_count"]
    assert runs >= 19 and births >= 30 and deaths >= 30


# ------------------------------------------------------------------------------------------ ProductMixture along a sweep
def product_mixture_matches_the_reference(abi, tol=2e-5):
    """src/product_mixture.cc compiled unmodified (tests/golden/reference_product_mixture.json: the score vector of
    ProductMixture::score_value -- score_diff on rows the reference's own Differ compressed -- before every ADD of a
    cat-kernel trajectory, then add_value / remove_value with the trajectory's packed groups) against the engine's own
    sweep of the same table with the debug taps on.  The trajectory is the engine's (same seed => same picks as when the
    fixture was made); what must agree at every one of the 830 ADDs is the reference's view of the state: how many
    groups there are, which packed index each has after every birth and death, the clustering prior and the sum over
    the observed features of every group.  Per-feature arithmetic on the reference side is the oracle's own (see
    oracle/ref_stubs_mix), so the tolerance only covers the reference's float32 accumulation."""
    import ctypes
    import mixture_pin_util as U
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_product_mixture.json")))
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
    lib.lo_py_score_counts.restype = ctypes.c_double
    lib.lo_py_score_counts.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
    adds = 0
    for case in fixture["cases"]:
        model, table = U.build(case)
        assert table == case["table"], case["name"]                     # the generator's table, rebuilt from the seed
        for mine, theirs in zip(model.features, case["features"]):
            assert mine == theirs, case["name"]
        tare = None
        if case["tare"]:
            obs = np.array([v is not None for v in case["tare"]], np.uint8)
            val = np.array([v or 0 for v in case["tare"]], np.uint32)
            tare = (obs, val)
        acts, picks, scores, score_data, counts = U.sweep(abi, case, model, table, tare)
        assert picks.tolist() == case["picks"], case["name"]
        assert len(scores) == len(case["adds"])
        for i, (mine, ref) in enumerate(zip(scores, case["adds"])):
            ref = np.asarray(ref, np.float64)
            assert len(mine) == len(ref), (case["name"], i)              # the same number of groups, empties included
            assert np.all(np.abs(mine - ref) <= tol * np.maximum(1.0, np.abs(ref))), (case["name"], i)
            adds += 1
        assert counts.tolist() == case["counts"], case["name"]           # packed order after the last action
        # CrossCat::score_data = the topology's score of the feature counts + every kind's mixture.score_data
        # (src/cross_cat.cc:238-250); the fixture holds the mixture's part
        one_kind = np.array([model.n_features], np.uint32)
        topology = lib.lo_py_score_counts(case["alpha"], case["d"], one_kind.ctypes.data, 1)
        assert abs(score_data - (case["score_data"] + topology)) <= 2e-5 * abs(score_data), case["name"]
    assert adds >= 800


def test_oracle_sweep_agrees_with_the_reference_product_mixture(oracle_abi):
    product_mixture_matches_the_reference(oracle_abi)


@pytest.mark.gpu
def test_cuda_sweep_agrees_with_the_reference_product_mixture(cuda_abi):
    product_mixture_matches_the_reference(cuda_abi, tol=4e-5)      # float32 tables on the device (1e-5 against the oracle) + the reference's float32 sums


# ------------------------------------------------------------------------------------------ a whole kind-structure pass
def test_oracle_kind_structure_pass_reproduces_the_reference(oracle_abi):
    """The reference's own CrossCat + KindKernel + KindProposer + ProductModel + ProductMixture + ValueSplitter +
    Assignments, compiled unmodified and run through a kind-structure pass (tests/golden/reference_kind_structure.json:
    KindKernel::add_row / remove_row for every action, try_run() per round, the kernel destroyed at the end; random
    decisions scripted from this oracle's run), against the oracle's init_featureless_kinds / cat_sweep /
    kind_proposer_score / kind_run on the same tables.  Must agree: the score vector of every one of the 1 670 draws
    over all kinds (ephemeral ones included), L[f][k] of every round -- the proposer's statistics are the FULL row added
    to every kind's group (a9) and every (feature, kind) pair is scored from them (a10) --, the feature -> kind map the
    block sampler and move_features leave (a11, a12), and at the end every surviving kind's counts and sufficient
    statistics (integers bit for bit: moved features carry the proposer's statistics, src/product_mixture.cc:886-915)
    and CrossCat::score_data (a14)."""
    import structure_pin_util as U
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_kind_structure.json")))
    vectors = rounds = 0
    for case in fixture["cases"]:
        model, table = U.build(case)
        mine = U.run(oracle_abi, case, model, table)
        vectors += U.compare(case, model, mine, case["reference"])
        rounds += case["n_rounds"]
    assert vectors >= 1600 and rounds >= 7


# ------------------------------------------------------------------------------------------ HyperKernel::run
def hyper_run_matches_the_reference(abi):
    """src/hyper_kernel.cc compiled unmodified (HyperKernel::run over infer_grid.hpp / hyper_prior.hpp and the real
    CrossCat: tests/golden/reference_hyper.json) against hyper_run on the same tables, groups, grids and uniforms: the
    grid point chosen for the topology, for every kind's clustering and for every coordinate of every feature (88 draws
    over 3 cases; single-point grids set without a draw) must be the reference's, which requires the same hypotheses
    scored in the same order from the same statistics with the same task -> random stream numbering
    (src/hyper_kernel.cc:195-235).  No DPD column."""
    import hyper_pin_util as U
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_hyper.json")))
    draws = 0
    for case in fixture["cases"]:
        model, table = U.build(case)
        U.compare(case, U.run(abi, case, model, table), case["reference"])
        draws += case["reference"]["draws"]
    assert draws >= 80


def test_oracle_hyper_run_reproduces_the_reference(oracle_abi):
    hyper_run_matches_the_reference(oracle_abi)


# ------------------------------------------------------------------------------------------ CatKernel over several kinds
def test_oracle_multi_kind_sweep_reproduces_the_reference_cat_kernel(oracle_abi):
    """src/cat_kernel.hpp compiled unmodified over the reference's own CrossCat, ValueSplitter, Assignments, ProductModel
    and ProductMixture -- and its own Differ and CrossCat::update_tares for the tared case
    (tests/golden/reference_cat_kernel.json: every action through CatKernel::add_row / remove_row(rng, row,
    assignments), draws scripted from this oracle's sweep) against the oracle's cat_sweep over the same multi-kind
    tables: the score vector of each of the 1 300 (row, kind) draws (which cells of a row reach which kind: a7), the
    final clustering counts of every kind, which rows are still assigned and the PACKED group of each -- the reference
    keeps global ids in FIFO queues and maps them through the id tracker (a8) --, and CrossCat::score_data (a14)."""
    import cat_pin_util as U
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_cat_kernel.json")))
    vectors = 0
    for case in fixture["cases"]:
        model, table = U.build(case)
        tare = None
        if case["tare_row"]:
            tare = (np.array([v is not None for v in case["tare_row"]], np.uint8), np.array([v or 0 for v in case["tare_row"]], np.uint32))
        vectors += U.compare(case, U.run(oracle_abi, case, model, table, tare), case["reference"])
    assert vectors >= 1300


# ------------------------------------------------------------------------------------------ the dump order of groups
def test_dump_order_of_groups_is_the_reference_s():
    """CrossCat::get_sorted_groupids compiled unmodified (tests/golden/reference_sorted_groupids.json) against
    lbio_sorted_groupids, the one implementation both hosts dump with (loom.hpp, loom_b200/tasks.py): non-empty groups,
    largest first, and groups of EQUAL count exactly where the reference's std::sort leaves them -- several of the
    fixture's kinds differ from what a stable sort would give."""
    from loom_b200 import pbio
    fixture = json.load(open(os.path.join(HERE, "golden", "reference_sorted_groupids.json")))
    unstable = 0
    for kind in fixture["kinds"]:
        counts = kind["counts"]
        assert pbio.sorted_groupids(counts).tolist() == kind["order"], len(counts)
        unstable += kind["order"] != sorted([g for g in range(len(counts)) if counts[g]], key=lambda g: -counts[g])
    assert len(fixture["kinds"]) >= 20 and unstable >= 5
