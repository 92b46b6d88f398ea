"""Generate the committed fixtures under tests/golden/.

The reference ships no numeric golden vectors for this path and cannot be built or
imported in this image, so nothing here comes from running loom's engine.  What is committed instead:

  closed_forms.json   known answers for every conjugate closed form on the hot path, computed by
                      scipy (an independent implementation): betabinom / nbinom / Student-t
                      predictives, gammaln marginals, the Pitman-Yor EPPF by sequential seating.
                      These pin the ORACLE's arithmetic (tests/test_golden.py, CPU).
  structural.json     the reference's own structural fixtures, transcribed with their source lines:
                      LATENT_SIZES / BELL_NUMBERS (loom/test/test_posterior_enum.py:65-79,778-781)
                      and the worked tare example (doc/adapting.md:323-414).
  reference_grids.json  outputs of the reference's own loom/gridding.py (the one piece of the
                      reference that runs here): pins loom_b200/hyperprior.py.
  engine_small.npz    a small mixed model + rows with the oracle's results on it (statistics for a
                      fixed assignment, row x group scores, data score, feature x kind marginals,
                      a seeded Gibbs sweep): regression vectors for the oracle (all of it) and for
                      the CUDA path (everything that does not depend on a sampled pick bit for
                      bit; the sweep itself is checked by replay in tests/test_cuda_cat.py).

  query_differ_small.npz  on the same model in its fixed-assignment state: the oracle's query scores, a seeded batch
                      of query draws, the tare row and a diff of the whole block (regression vectors for the
                      oracle; the CUDA path reproduces the scores to 1e-5, the diff bit for bit and the draws
                      wherever its group pick agrees).

Run from the repo root:  python tests/golden/make_golden.py   (or `... make_golden.py query_differ_small`)
"""
import json
import os
import sys

import numpy as np
from scipy import special, stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def f32(x):
    return float(np.float32(x))


def closed_forms(rng):
    cases = []
    for _ in range(12):   # BB
        a, b = f32(rng.uniform(0.1, 4)), f32(rng.uniform(0.1, 4))
        h, t = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        cases.append({"type": "bb", "hp": [a, b], "ints": [h, t], "flts": [],
                      "values": [0, 1],
                      "predictive": [float(stats.betabinom.logpmf(x, 1, a + h, b + t)) for x in (0, 1)],
                      "marginal": float(special.betaln(a + h, b + t) - special.betaln(a, b))})
    for dim in (2, 4, 16, 256):   # DD
        for _ in range(3):
            al = np.float32(rng.uniform(0.05, 3, dim)).astype(np.float64)
            n = rng.integers(0, 30, dim) * (rng.random(dim) < 0.6)
            vals = [0, dim // 2, dim - 1]
            cases.append({"type": "dd", "dim": dim, "aux": al.tolist(), "ints": n.tolist() + [int(n.sum())], "flts": [],
                          "values": vals,
                          "predictive": [float(np.log((al[v] + n[v]) / (al.sum() + n.sum()))) for v in vals],
                          "marginal": float((special.gammaln(al + n) - special.gammaln(al)).sum()
                                            + special.gammaln(al.sum()) - special.gammaln(al.sum() + n.sum()))})
    for dim in (3, 8):   # DPD (dense ids; betas sum with beta0 to 1)
        for _ in range(3):
            w = rng.dirichlet(np.ones(dim + 1))
            betas = np.float32(w[:dim]).astype(np.float64)
            beta0 = f32(w[dim])
            alpha, gamma = f32(rng.uniform(0.3, 5)), f32(rng.uniform(0.3, 5))
            n = rng.integers(0, 20, dim)
            vals = list(range(dim))[:3]
            ab = alpha * betas
            cases.append({"type": "dpd", "dim": dim, "hp": [gamma, alpha, beta0], "aux": betas.tolist(),
                          "ints": n.tolist() + [int(n.sum())], "flts": [], "values": vals,
                          "predictive": [float(np.log((ab[v] + n[v]) / (alpha + n.sum()))) for v in vals],
                          "marginal": float((special.gammaln(ab + n) - special.gammaln(ab)).sum()
                                            + special.gammaln(alpha) - special.gammaln(alpha + n.sum()))})
    for _ in range(12):   # GP
        al, ib = f32(rng.uniform(0.2, 5)), f32(rng.uniform(0.1, 3))
        xs = rng.poisson(rng.uniform(0.5, 30), int(rng.integers(0, 40)))
        n, S = len(xs), int(xs.sum())
        log_prod = float(special.gammaln(xs + 1.0).sum())
        a, b = al + S, ib + n
        vals = [0, 1, 7, 50]
        # marginal: chain rule of negative-binomial predictives
        m, cn, cs = 0.0, 0, 0
        for x in xs:
            m += float(stats.nbinom.logpmf(x, al + cs, (ib + cn) / (ib + cn + 1)))
            cn, cs = cn + 1, cs + int(x)
        cases.append({"type": "gp", "hp": [al, ib], "ints": [n, S], "flts": [log_prod], "values": vals,
                      "predictive": [float(stats.nbinom.logpmf(x, a, b / (b + 1))) for x in vals],
                      "marginal": m})
    for _ in range(12):   # NICH
        mu, kappa = f32(rng.normal(0, 2)), f32(rng.uniform(0.2, 3))
        sigmasq, nu = f32(rng.uniform(0.2, 4)), f32(rng.uniform(0.5, 6))
        xs = np.float32(rng.normal(rng.normal(0, 3), rng.uniform(0.3, 3), int(rng.integers(0, 40)))).astype(np.float64)
        n = len(xs)
        mean = float(xs.mean()) if n else 0.0
        ctv = float(((xs - mean) ** 2).sum()) if n else 0.0

        def post(n, mean, ctv):
            kn, nun = kappa + n, nu + n
            mun = (kappa * mu + n * mean) / kn
            sn = (nu * sigmasq + ctv + n * kappa * (mu - mean) ** 2 / kn) / nun
            return nun, mun, np.sqrt((1 + kn) * sn / kn)
        nun, mun, scale = post(n, mean, ctv)
        vals = [f32(v) for v in (-4.0, 0.0, 0.37, 12.5)]
        m = 0.0
        for i, x in enumerate(xs):   # chain rule of Student-t predictives
            pm = float(xs[:i].mean()) if i else 0.0
            pc = float(((xs[:i] - pm) ** 2).sum()) if i else 0.0
            d, l, s = post(i, pm, pc)
            m += float(stats.t.logpdf(x, d, loc=l, scale=s))
        cases.append({"type": "nich", "hp": [mu, kappa, sigmasq, nu], "ints": [n], "flts": [mean, ctv],
                      "values": [int(np.float32(v).view(np.uint32)) for v in vals],
                      "predictive": [float(stats.t.logpdf(v, nun, loc=mun, scale=scale)) for v in vals],
                      "marginal": m})
    py = []
    for _ in range(10):   # Pitman-Yor EPPF by sequential seating
        alpha, d = f32(rng.uniform(0.2, 10)), f32(rng.uniform(0.0, 0.8))
        counts = [int(c) for c in rng.integers(0, 9, int(rng.integers(1, 9)))]
        logp, n, m = 0.0, 0, 0
        for c in counts:
            if not c:
                continue
            logp += np.log((alpha + d * m) / (n + alpha))
            n += 1
            for j in range(1, c):
                logp += np.log((j - d) / (n + alpha))
                n += 1
            m += 1
        py.append({"alpha": alpha, "d": d, "counts": counts, "score_counts": float(logp)})
    return {"source": "scipy %s" % __import__("scipy").__version__, "features": cases, "pitman_yor": py}


def structural():
    import pe_util
    return {
        "LATENT_SIZES": {"source": "loom/test/test_posterior_enum.py:65-79", "value": pe_util.LATENT_SIZES},
        "BELL_NUMBERS": {"source": "loom/test/test_posterior_enum.py:778-781", "value": pe_util.BELL_NUMBERS},
        "tare_example": {
            "source": "doc/adapting.md:323-414",
            "row": [False, True, None, False, False],
            "tare": {"observed": [1, 1, 0, 0, 0], "booleans": [False, False]},
            "diff": {"pos": {"observed": [0, 1, 0, 1, 1], "booleans": [True, False, False]},
                     "neg": {"observed": [0, 1, 0, 0, 0], "booleans": [False]},
                     "tares": [0]},
        },
    }


def engine_small():
    from loom_b200 import Engine, synth, sweep_actions
    from loom_b200._abi import Abi
    oracle = Abi(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "libloom_oracle.so"), "lo_")
    layout = [("bb", 3), ("dd4", 2), ("dd16", 2), ("dd256", 1), ("dpd8", 2), ("gp", 3), ("nich", 3)]
    f2k = [0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0]
    model, rows, truth = synth.generate(600, layout, density=0.7, seed=2026, f2k=f2k)
    out = {"layout": json.dumps(layout), "n_rows": rows.n_rows, "cat8": rows.cat8, "wide": rows.wide, "mask": rows.mask,
           "f2k": model.f2k, "features": json.dumps(model.features), "kind_py": model.kind_py}
    e = Engine(oracle)
    e.set_model(model)
    e.set_rows(rows)
    # (1) statistics of a fixed assignment (K2), scores against them (K1), data score, L[f][k] (K4+K5)
    for k in range(model.n_kinds):
        a = truth["assign"][k].astype(np.uint32)
        uniq, inv = np.unique(a, return_inverse=True)
        out["fixed_assign_%d" % k] = inv.astype(np.uint32)
        e.set_groups(k, len(uniq))
        e.set_assignments(k, inv.astype(np.uint32))
    e.accumulate_rows()
    for k in range(model.n_kinds):
        c, i, f = e.get_groups(k)
        out["fixed_counts_%d" % k], out["fixed_ints_%d" % k], out["fixed_flts_%d" % k] = c, i, f
        out["fixed_scores_%d" % k] = e.score_rows(k)[:64].astype(np.float64)
    out["fixed_score_data"] = np.float64(e.score_data())
    out["fixed_L"] = e.kind_proposer_score().astype(np.float64)
    # (2) a seeded sweep from empty: greedy ADD pass then one REMOVE/ADD pass
    e2 = Engine(oracle)
    e2.set_model(model)
    e2.set_rows(rows)
    acts = np.concatenate([sweep_actions(rows.n_rows), sweep_actions(rows.n_rows, add=True, remove=True)])
    out["sweep_actions"] = acts
    out["sweep_seed"] = np.uint64(77)
    picks, scores = e2.cat_sweep(acts, seed=77, debug=True, debug_stride=8)
    out["sweep_picks"] = picks
    for k in range(model.n_kinds):
        out["sweep_assign_%d" % k] = e2.get_assignments(k)
        c, i, f = e2.get_groups(k)
        out["sweep_counts_%d" % k], out["sweep_ints_%d" % k] = c, i
    out["sweep_score_data"] = np.float64(e2.score_data())
    return out


def query_differ_small():
    """The rows SURVEY.md 8f added (query n3, tare / sparsify n4) on the engine_small model in its fixed-assignment
    state: Score of 64 rows, 256 joint draws given row 0, the tare row and the diff of the whole block."""
    from loom_b200 import Engine, Model, RowBlock
    from loom_b200._abi import Abi
    oracle = Abi(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "libloom_oracle.so"), "lo_")
    z = np.load(os.path.join(HERE, "engine_small.npz"))
    model = Model(json.loads(str(z["features"])), z["f2k"], kind_py=z["kind_py"])
    rows = RowBlock(int(z["n_rows"]), z["cat8"], z["wide"], z["mask"])
    e = Engine(oracle)
    e.set_model(model)
    e.set_rows(rows)
    for k in range(model.n_kinds):
        a = z["fixed_assign_%d" % k]
        e.set_groups(k, int(a.max()) + 1)
        e.set_assignments(k, a)
    e.accumulate_rows()
    out = {"query_scores": e.query_score(0, 64).astype(np.float64)}
    to_sample = np.arange(1, model.n_features, 2, dtype=np.uint32)
    values, groups = e.query_sample(0, to_sample, 256, seed=5, groups=True)
    out.update(to_sample=to_sample, sample_seed=np.uint64(5), sample_values=values, sample_groups=groups)
    obs, val = e.make_tare()
    # engine_small has no dominant values; a hand-made tare exercises every branch of the diff encoding
    hand_obs = np.zeros(model.n_features, np.uint8)
    hand_val = np.zeros(model.n_features, np.uint32)
    for f, feat in enumerate(model.features):
        if feat["type"] != "nich" and f % 2 == 0:
            hand_obs[f], hand_val[f] = 1, 1
    d = e.sparsify(hand_obs, hand_val)
    out.update(tare_observed=obs, tare_values=val, hand_tare_observed=hand_obs, hand_tare_values=hand_val,
               **{k: d[k] for k in ("pos_ptr", "pos_feature", "pos_value", "neg_ptr", "neg_feature")})
    return out


def reference_grids():
    """Outputs of the reference's own loom/gridding.py (pure numpy, importable by file path even
    though the loom package is not): pins loom_b200/hyperprior.py.  Only runs where
    /root/reference exists; the JSON it writes is what travels."""
    import importlib.util
    path = "/root/reference/loom/gridding.py"
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_gridding", path)
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    return {"source": "loom/gridding.py executed from /root/reference",
            "uniform_0_1_7": g.uniform(0, 1, 7).tolist(),
            "center_heavy_0_1_20": g.center_heavy(0, 1, 20).tolist(),
            "left_heavy_m1_2_30": g.left_heavy(-1, 2, 30).tolist(),
            "right_heavy_m1_1_20": g.right_heavy(-1, 1, 20).tolist(),
            "pitman_yor": g.pitman_yor()}


if __name__ == "__main__":
    if sys.argv[1:] == ["query_differ_small"]:      # added later: leaves the other fixtures untouched
        np.savez_compressed(os.path.join(HERE, "query_differ_small.npz"), **query_differ_small())
        sys.exit(0)
    rng = np.random.default_rng(20261019)
    with open(os.path.join(HERE, "closed_forms.json"), "w") as f:
        json.dump(closed_forms(rng), f, indent=0)
    with open(os.path.join(HERE, "structural.json"), "w") as f:
        json.dump(structural(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "engine_small.npz"), **engine_small())
    np.savez_compressed(os.path.join(HERE, "query_differ_small.npz"), **query_differ_small())
    grids = reference_grids()
    if grids is not None:
        with open(os.path.join(HERE, "reference_grids.json"), "w") as f:
            json.dump(grids, f, indent=0)
    print("wrote", sorted(os.listdir(HERE)))
