"""Pin the oracle's arithmetic against independent known answers (scipy closed
forms, Random123 Philox vectors).  The reference ships no numeric golden
vectors, so these are the oracle's anchors."""
import ctypes as C

import numpy as np
import pytest
from scipy import special, stats

from loom_b200._abi import FeatureSpec, LB_BB, LB_DD, LB_DPD, LB_GP, LB_NICH


def _spec(t, dim=0, hp=(), aux_off=0):
    s = FeatureSpec()
    s.type, s.dim, s.aux_off = t, dim, aux_off
    for i, v in enumerate(hp):
        s.hp[i] = v
    return s


def _bind(oracle_abi):
    lib = oracle_abi.lib
    lib.lo_feature_score_value.restype = C.c_double
    lib.lo_feature_score_value.argtypes = [C.POINTER(FeatureSpec), C.POINTER(C.c_float),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_uint32]
    lib.lo_feature_score_data.restype = C.c_double
    lib.lo_feature_score_data.argtypes = lib.lo_feature_score_value.argtypes[:4]
    lib.lo_py_score_counts.restype = C.c_double
    lib.lo_py_score_counts.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_uint32), C.c_int]
    lib.lo_uniform.restype = C.c_double
    lib.lo_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
    return lib


def _call(fn, spec, aux, ints, flts, *rest):
    aux = np.asarray(aux if len(aux) else [0], np.float32)
    ints = np.asarray(ints, np.uint32)
    flts = np.asarray(flts if len(flts) else [0], np.float64)
    return fn(C.byref(spec), aux.ctypes.data_as(C.POINTER(C.c_float)),
              ints.ctypes.data_as(C.POINTER(C.c_uint32)),
              flts.ctypes.data_as(C.POINTER(C.c_double)), *rest)


def test_philox_known_answer(oracle_abi):
    # Random123 kat_vectors: philox4x32-10
    lib = _bind(oracle_abi)
    f = lib.lo_philox4x32
    f.argtypes = [C.POINTER(C.c_uint32)] * 3
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        f(c, k, o)
        assert list(o) == want
    u = [lib.lo_uniform(7, 0, i, 3) for i in range(2000)]
    assert 0 <= min(u) and max(u) < 1 and abs(np.mean(u) - 0.5) < 0.03


def test_bb_matches_betabinom(oracle_abi):
    lib = _bind(oracle_abi)
    a, b, h, t = float(np.float32(0.7)), 2.5, 11, 4
    s = _spec(LB_BB, hp=(a, b))
    p1 = np.exp(_call(lib.lo_feature_score_value, s, [], [h, t], [], 1))
    assert p1 == pytest.approx((a + h) / (a + b + h + t), rel=1e-12)
    # marginal of an ordered sequence with h heads, t tails
    want = special.betaln(a + h, b + t) - special.betaln(a, b)
    assert _call(lib.lo_feature_score_data, s, [], [h, t], []) == pytest.approx(want, rel=1e-12)


def test_dd_matches_dirichlet_multinomial(oracle_abi):
    lib = _bind(oracle_abi)
    al32 = np.float32([0.5, 1.5, 2.0, 0.25])
    al = al32.astype(np.float64)
    n = [3, 0, 7, 1]
    s = _spec(LB_DD, dim=4)
    ints = n + [sum(n)]
    for v in range(4):
        got = _call(lib.lo_feature_score_value, s, al, ints, [], v)
        assert got == pytest.approx(np.log((al[v] + n[v]) / (al.sum() + sum(n))), rel=1e-12)
    want = (special.gammaln(al + n) - special.gammaln(al)).sum() + special.gammaln(al.sum()) \
        - special.gammaln(al.sum() + sum(n))
    assert _call(lib.lo_feature_score_data, s, al, ints, []) == pytest.approx(want, rel=1e-12)
    # chain rule: marginal == sum of sequential predictives
    seq, cur, total = [0, 0, 0, 2, 2, 2, 2, 2, 2, 2, 3], [0] * 5, 0.0
    for v in seq:
        total += _call(lib.lo_feature_score_value, s, al, cur, [], v)
        cur[v] += 1
        cur[4] += 1
    assert total == pytest.approx(want, rel=1e-12)


def test_dpd_is_dd_with_alpha_beta(oracle_abi):
    lib = _bind(oracle_abi)
    betas = np.float32([0.2, 0.3, 0.1]).astype(np.float64)
    alpha = float(np.float32(1.7))
    n = [2, 0, 5]
    s = _spec(LB_DPD, dim=3, hp=(1.0, alpha, 0.4))
    got = _call(lib.lo_feature_score_value, s, betas, n + [7], [], 2)
    assert got == pytest.approx(np.log((alpha * betas[2] + 5) / (alpha + 7)), rel=1e-12)


def test_gp_matches_negative_binomial(oracle_abi):
    lib = _bind(oracle_abi)
    al, ib = 2.5, 0.75
    xs = [3, 0, 9, 4, 4]
    n, S = len(xs), sum(xs)
    log_prod = float(sum(special.gammaln(x + 1) for x in xs))
    s = _spec(LB_GP, hp=(al, ib))
    a, b = al + S, ib + n
    for x in (0, 1, 6, 40):
        got = _call(lib.lo_feature_score_value, s, [], [n, S], [log_prod], x)
        assert got == pytest.approx(stats.nbinom.logpmf(x, a, b / (b + 1)), rel=1e-10)
    # marginal by the chain rule of predictives
    total, cn, cs, cl = 0.0, 0, 0, 0.0
    for x in xs:
        total += _call(lib.lo_feature_score_value, s, [], [cn, cs], [cl], x)
        cn, cs, cl = cn + 1, cs + x, cl + float(special.gammaln(x + 1))
    assert _call(lib.lo_feature_score_data, s, [], [n, S], [log_prod]) == pytest.approx(total, rel=1e-10)


def test_nich_matches_student_t(oracle_abi):
    lib = _bind(oracle_abi)
    mu, kappa, sigmasq, nu = float(np.float32(0.3)), 1.5, 2.0, 3.0
    xs = np.float32([0.1, -1.2, 2.5, 0.7])
    n, mean = len(xs), float(np.mean(xs.astype(np.float64)))
    ctv = float(((xs.astype(np.float64) - mean) ** 2).sum())
    s = _spec(LB_NICH, hp=(mu, kappa, sigmasq, nu))
    kn, nun = kappa + n, nu + n
    mun = (kappa * mu + n * mean) / kn
    sn = (nu * sigmasq + ctv + n * kappa * (mu - mean) ** 2 / kn) / nun
    scale = np.sqrt((1 + kn) * sn / kn)
    for x in (-3.0, 0.0, 0.5, 10.0):
        raw = int(np.float32(x).view(np.uint32))
        got = _call(lib.lo_feature_score_value, s, [], [n], [mean, ctv], raw)
        assert got == pytest.approx(stats.t.logpdf(x, nun, loc=mun, scale=scale), rel=1e-10)
    total, cn, cm, cc = 0.0, 0, 0.0, 0.0
    for x in xs.astype(np.float64):
        raw = int(np.float32(x).view(np.uint32))
        total += _call(lib.lo_feature_score_value, s, [], [cn], [cm, cc], raw)
        cn += 1
        d = x - cm
        cm += d / cn
        cc += d * (x - cm)
    assert _call(lib.lo_feature_score_data, s, [], [n], [mean, ctv]) == pytest.approx(total, rel=1e-10)


def test_pitman_yor_score_counts_is_eppf(oracle_abi):
    lib = _bind(oracle_abi)
    alpha, d = 2.0, 0.1
    counts = np.uint32([3, 1, 0, 4])
    got = lib.lo_py_score_counts(alpha, d, counts.ctypes.data_as(C.POINTER(C.c_uint32)), 4)
    # sequential seating probability of the partition {3,1,4}
    logp, n, m = 0.0, 0, 0
    for c in (3, 1, 4):
        logp += np.log((alpha + d * m) / (n + alpha))
        n += 1
        for j in range(1, c):
            logp += np.log((j - d) / (n + alpha))
            n += 1
        m += 1
    assert got == pytest.approx(logp, rel=1e-12)
