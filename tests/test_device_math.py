"""The device's score arithmetic on the CPU suite: loom_b200/csrc/lb_common.cuh (lb_refresh_row + lb_score_cell, the
float32 formulas k_cat_sweep* and k_score_rows run) compiled for the host (tests/cpp/device_math_host.cc through
tests/cpp/cuda_shim/) against the float64 oracle's per-feature closed forms, for RANDOM hyperparameters and group sizes
from empty to 10^6 rows.  The GPU parity tests visit the default hypers (alpha = 2, 0.5, ...) almost exclusively; after
hyper inference the hypers are arbitrary grid values, which is where a float32 formula loses digits first (the oracle's
own Gamma-Poisson cache had such a loss: see DESIGN.md section 2).  Host libm is not CUDA's, so this is an accuracy
bound on the formulas, not a bit-level check."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from loom_b200.engine import FeatureSpec

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TOL = 1e-5                      # the score tolerance of the GPU parity tests (tests/test_cuda_cat.py SCORE_TOL)


@pytest.fixture(scope="module")
def device_math(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("device_math") / "device_math_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-I", os.path.join(HERE, "cpp", "cuda_shim"), "-o", exe,
                           os.path.join(HERE, "cpp", "device_math_host.cc")])
    return exe


def _cases(rng, n):
    """(type, dim, hp[4], aux, ints, flts, raw): sufficient statistics of ONE group of `size` rows"""
    out = []
    for _ in range(n):
        t = int(rng.integers(0, 5))
        size = int(rng.choice([0, 1, 3, 40, 1000, 30000, 10 ** 6]))
        hp, aux, ints, flts, dim = [0.0] * 4, [], [], [], 0
        if t == 0:
            hp[:2] = 10 ** rng.uniform(-1, 1, 2)
            heads = int(rng.binomial(size, rng.uniform(0.01, 0.99))) if size else 0
            ints, raw = [heads, size - heads], int(rng.integers(0, 2))
        elif t in (1, 2):
            dim = int(rng.choice([2, 5, 16, 256]))
            if t == 1:
                aux = (10 ** rng.uniform(-1, 1, dim)).tolist()
            else:
                hp[:3] = [10 ** rng.uniform(-1, 2), 10 ** rng.uniform(-1, 1), 0.05]
                b = rng.dirichlet(np.ones(dim)) * 0.95
                aux = np.maximum(b, 1e-4).tolist()
            counts = rng.multinomial(size, rng.dirichlet(np.ones(dim) * 0.3)) if size else np.zeros(dim, int)
            ints, raw = counts.tolist() + [size], int(rng.integers(0, dim))
        elif t == 3:
            hp[:2] = [10 ** rng.uniform(-1, 2), 10 ** rng.uniform(-2, 1)]
            lam = 10 ** rng.uniform(-1, 2)
            ints = [size, int(rng.poisson(lam * size)) if size else 0]
            flts, raw = [0.0], int(rng.poisson(lam) if rng.uniform() < 0.8 else rng.poisson(5 * lam + 3))
        else:
            hp[:4] = [rng.uniform(-3, 3), 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(0, 1.5)]
            mean, sd = rng.uniform(-50, 50), 10 ** rng.uniform(-1, 1.5)
            ints, flts = [size], ([mean, sd * sd * size] if size else [0.0, 0.0])
            raw = int(np.float32(mean + sd * rng.standard_normal() * rng.choice([1, 4])).view(np.uint32))
        hp = [float(np.float32(h)) for h in hp]
        aux = [float(np.float32(a)) for a in aux]
        out.append((t, dim, hp, aux, ints, flts, raw))
    return out


def test_device_score_formulas_against_the_float64_oracle(device_math):
    lib = C.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
    lib.lo_feature_score_value.restype = C.c_double
    rng = np.random.default_rng(7)
    cases = _cases(rng, 4000)
    text = ""
    for t, dim, hp, aux, ints, flts, raw in cases:
        text += "%d %d %s %d %s %d %s %d %s %d\n" % (t, dim, " ".join("%.9g" % h for h in hp), len(aux), " ".join("%.9g" % a for a in aux),
                                                  len(ints), " ".join(str(i) for i in ints), len(flts), " ".join("%.17g" % x for x in flts), raw)
    got = np.array(subprocess.run([device_math], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.split(), np.float64)
    assert len(got) == len(cases)
    # err = |device - oracle| / max(1, |oracle|), the metric of the GPU parity tests, per regime
    envelope = {}
    for (t, dim, hp, aux, ints, flts, raw), mine in zip(cases, got):
        spec = FeatureSpec()
        spec.type, spec.dim, spec.aux_off = t, dim, 0
        for j in range(4):
            spec.hp[j] = hp[j]
        a = (C.c_float * max(len(aux), 1))(*aux)
        i = (C.c_uint32 * max(len(ints), 1))(*ints)
        f = (C.c_double * max(len(flts), 2))(*flts)
        want = lib.lo_feature_score_value(C.byref(spec), a, i, f, C.c_uint32(raw))
        err = abs(mine - want) / max(1.0, abs(want))
        size = ints[0] + ints[1] if t == 0 else ints[-1] if t in (1, 2) else ints[0]
        if t == 3:
            regime = "gp, value <= 20" if raw <= 20 and ints[1] < 2 ** 24 else "gp, value > 20"
        elif t == 4:
            regime = "nich, group <= 1000 rows" if size <= 1000 else "nich, group > 1000 rows"
        else:
            regime = ["bb", "dd", "dpd"][t]
        if err > envelope.get(regime, 0.0) and os.environ.get("LB_DEVICE_MATH_VERBOSE"):
            print(regime, "%.2e" % err, mine, want, hp[:2], ints, raw)
        envelope[regime] = max(envelope.get(regime, 0.0), err)
    print({k: "%.1e" % v for k, v in sorted(envelope.items())})
    # Bernoulli / categorical tables: one log per entry, nothing cancels
    for regime in ("bb", "dd", "dpd"):
        assert envelope[regime] <= 2e-6, (regime, envelope[regime])
    # the Student-t predictive holds 1e-5 everywhere since the posterior mean is formed with a float64 division (a float32
    # reciprocal there cost 1.7e-5 at 10^6-row groups whose |mean| is hundreds of sigmas: found with this test)
    assert envelope["nich, group <= 1000 rows"] <= 3e-6, envelope
    assert envelope["nich, group > 1000 rows"] <= 3e-6, envelope
    # Gamma-Poisson: 1e-5 in both regimes since counts above 8 take the combined-Stirling form (lb_gp_score_stirling); the
    # textbook sum was 5e-5 off for counts in the hundreds
    assert envelope["gp, value <= 20"] <= TOL, envelope
    assert envelope["gp, value > 20"] <= TOL, envelope


def test_gamma_poisson_score_over_extreme_regimes(device_math):
    """Gamma-Poisson alone, far outside what data looks like: groups of 1 to 3 10^6 rows, means from 0.03 to 1000, alpha
    and inv_beta over four decades, values at the mean, far in the tail (scores of -10^4) and uniform up to 2000.  What
    float32 can hold here is bounded by a * ulp of the O(1e-4) quantities the form combines: 99 % of the cases inside
    1e-5, all inside 6e-5 (the textbook sum: 99 % inside 1e-4, worst 6e-4 on a score of -4.5)."""
    lib = C.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
    lib.lo_feature_score_value.restype = C.c_double
    rng = np.random.default_rng(123)
    cases = []
    for _ in range(20000):
        size = int(10 ** rng.uniform(0, 6.5))
        lam = 10 ** rng.uniform(-1.5, 3)
        alpha, inv_beta = float(np.float32(10 ** rng.uniform(-2, 2.5))), float(np.float32(10 ** rng.uniform(-3, 1.5)))
        total = int(min(rng.poisson(lam * size), 4e9))
        mode = rng.integers(0, 4)
        x = int(rng.poisson(lam)) if mode == 0 else int(rng.poisson(lam * 3 + 5)) if mode == 1 else int(rng.integers(0, 2000)) if mode == 2 \
            else int(rng.integers(7, 12))
        cases.append((alpha, inv_beta, size, total, x))
    text = "".join("3 0 %.9g %.9g 0 0 0 2 %d %d 1 0.0 %d\n" % c for c in cases)
    got = np.array(subprocess.run([device_math], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.split(), np.float64)
    assert len(got) == len(cases) and np.isfinite(got).all()
    errs = np.empty(len(cases))
    for j, ((alpha, inv_beta, size, total, x), mine) in enumerate(zip(cases, got)):
        spec = FeatureSpec()
        spec.type, spec.hp[0], spec.hp[1] = 3, alpha, inv_beta
        want = lib.lo_feature_score_value(C.byref(spec), (C.c_float * 1)(0), (C.c_uint32 * 2)(size, total), (C.c_double * 2)(0, 0), C.c_uint32(x))
        errs[j] = abs(mine - want) / max(1.0, abs(want))
    assert np.quantile(errs, 0.99) <= 1e-5 and errs.max() <= 6e-5, (np.quantile(errs, [0.5, 0.99, 0.999]), errs.max())


@pytest.fixture(scope="module")
def device_marginal(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("device_marginal") / "device_marginal_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-I", os.path.join(HERE, "cpp", "cuda_shim"), "-o", exe,
                           os.path.join(HERE, "cpp", "device_marginal_host.cc")])
    return exe


def test_device_marginal_formulas_against_the_float64_oracle(device_marginal):
    """lb_marginal.cuh (lb_marginal_group: K5's scalar path and K7; lb_lgamma_rise: K5's Dirichlet table walk) against
    the oracle's om_score_data over the same random (hypers, group) pairs as the score test, per group.  What matters
    downstream is DIFFERENCES of sums of these over the groups of a kind (a feature's score under one kind against
    another, one grid point against the next), so the bound is relative to the marginal itself, as in the GPU tests."""
    lib = C.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
    lib.lo_feature_score_data.restype = C.c_double
    rng = np.random.default_rng(11)
    cases = [c for c in _cases(rng, 4000)]
    text = ""
    for t, dim, hp, aux, ints, flts, raw in cases:
        if t == 3:                                          # log_prod consistent with (count, sum): count cells of the mean value
            flts = [float(ints[0]) * float(np.log(np.arange(1, max(ints[1] // max(ints[0], 1), 1) + 1)).sum())]
        text += "%d %d %s %d %s %d %s %d %s\n" % (t, dim, " ".join("%.9g" % h for h in hp), len(aux), " ".join("%.9g" % a for a in aux),
                                               len(ints), " ".join(str(i) for i in ints), len(flts), " ".join("%.17g" % x for x in flts))
    got = np.array(subprocess.run([device_marginal], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.split(),
                   np.float64).reshape(-1, 2)
    assert len(got) == len(cases) and np.isfinite(got).all()
    # A log marginal of an n-row group is a few lgamma differences of size n ln n that largely cancel.  Taken in float32 each
    # is off by an ulp of its own size -- 0.8 nat at 10^6 rows, 140 nats for a Gamma-Poisson group whose values sum to 10^8,
    # the precision class of the reference's float32 arithmetic -- however small the marginal itself is.
    # Bound: 1e-5 of the marginal plus ulp * n ln(n + 3) (n = rows, or the sum of the values for Gamma-Poisson).
    envelope, slack = {}, {}
    for (t, dim, hp, aux, ints, flts, raw), (group, table) in zip(cases, got):
        if t == 3:
            flts = [float(ints[0]) * float(np.log(np.arange(1, max(ints[1] // max(ints[0], 1), 1) + 1)).sum())]
        spec = FeatureSpec()
        spec.type, spec.dim, spec.aux_off = t, dim, 0
        for j in range(4):
            spec.hp[j] = hp[j]
        a = (C.c_float * max(len(aux), 1))(*aux)
        i = (C.c_uint32 * max(len(ints), 1))(*ints)
        f = (C.c_double * max(len(flts), 2))(*flts)
        want = lib.lo_feature_score_data(C.byref(spec), a, i, f)
        n = ints[0] + ints[1] if t == 0 else ints[-1] if t in (1, 2) else ints[0]
        big = max(n, ints[1]) if t == 3 else n                  # Gamma-Poisson: the lgamma difference runs over the SUM of the values
        name = ["bb", "dd", "dpd", "gp", "nich"][t]
        for label, mine in ((name, group),) + (((name + " (table walk)", table),) if t in (1, 2) else ()):
            # lb_marginal_group takes differences over more than 4096 counts in float64 (lb_lgamma_diff_big): what is left is
            # the float32 difference below that, summed over up to 256 categories; K5's Dirichlet table walk stays float32
            ulp = 2e-7 if label.endswith("(table walk)") else 1e-7
            allowed = TOL * max(1.0, abs(want)) + ulp * big * np.log(big + 3.0)
            envelope[label] = max(envelope.get(label, 0.0), abs(mine - want))
            slack[label] = max(slack.get(label, 0.0), abs(mine - want) / allowed)
    print("largest absolute error (nats):", {k: "%.2g" % v for k, v in sorted(envelope.items())})
    print("largest error / allowed:", {k: "%.2f" % v for k, v in sorted(slack.items())})
    for label, ratio in slack.items():
        assert ratio <= 1.0, (label, slack, envelope)
    # in nats, over groups of up to 10^6 rows (sums of up to 10^8): what the float64 path bought
    assert envelope["bb"] <= 0.01 and envelope["gp"] <= 0.01 and envelope["nich"] <= 0.01, envelope      # were 1.7, 140, 0.64
    assert envelope["dd"] <= 0.1 and envelope["dpd"] <= 0.1, envelope                                    # were 1.8, 2.2
