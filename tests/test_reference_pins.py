"""Pins of the oracle against pieces of the REFERENCE ITSELF compiled in the build container (oracle/Makefile `ref`,
oracle/_ref/): the committed fixtures tests/golden/reference_*.json hold what the reference's own code printed, the
generators sit beside them.  (The schedule classes are pinned the same way in tests/test_infer_driver.py.)"""
import ctypes
import json
import os

import numpy as np

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
