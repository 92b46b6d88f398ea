"""Synthetic CrossCat inputs: a restatement of `loom.generate`.

Follows loom/generate.py:63-135 (exponential kind structure, one PY(2.0, 0.1)
clustering per kind) and src/generate.hpp:36-93 (per kind: draw the row's group
from the CRP, Bernoulli(density) observation mask, sample each observed value
from the group's predictive).  Values are drawn by first sampling each group's
latent parameters from the feature's prior and then i.i.d. values -- the same
distribution as the reference's sequential posterior-predictive sampling.
Upstream `EXAMPLES` hypers are unavailable (distributions is not vendored), so
hypers are the documented constants in `DEFAULT_HYPERS`.
Everything is numpy `default_rng(seed)`; nothing here touches oracle/.
"""
import numpy as np

from .engine import Model, RowBlock, pack_mask

CLUSTERING = (2.0, 0.1)  # loom/generate.py:41

DEFAULT_HYPERS = {
    "bb": {"alpha": 0.5, "beta": 0.5},
    "dd_alpha": 0.5,
    "dpd": {"gamma": 1.0, "alpha": 1.0},
    "gp": {"alpha": 2.0, "inv_beta": 0.5},
    "nich": {"mu": 0.0, "kappa": 1.0, "sigmasq": 1.0, "nu": 2.0},
}

# BASELINE.json configs restated as concrete shapes.
CONFIGS = {
    "C1": dict(rows=1000, types=[("bb", 10)], density=0.5, seed=1),
    "C2": dict(rows=100_000, types=[("bb", 25), ("dd256", 25), ("gp", 25), ("nich", 25)],
               density=0.5, seed=2),
    "C3": dict(rows=1_000_000, types=[("dd256", 256)], density=0.5, seed=3),
    "C3-16": dict(rows=1_000_000, types=[("dd16", 256)], density=0.5, seed=3),
    # taxi-like (examples/taxi/schema.json proportions): 200 dense columns (10 dictionary-encoded DPD, 100 DD of mixed
    # cardinality, 90 NICH) observed with probability 0.9 + 300 sparse booleans, always observed, on with probability
    # ~0.02 -- the columns loom_tare turns into one tare row
    "C4": dict(rows=10_000_000, types=[("sbb", 300), ("dd4", 30), ("dd16", 40), ("dd256", 30), ("dpd32", 10), ("nich", 90)],
               density=0.9, seed=4),
    # C5: 100M x 1000, a quarter each BB / DD-256 / GP / NICH; kinds sharded over 8 GPUs for the sweep, rows for the
    # kind and hyper kernels (runs at N = 8 only, in windows)
    "C5": dict(rows=100_000_000, types=[("bb", 250), ("dd256", 250), ("gp", 250), ("nich", 250)], density=0.5, seed=5),
}


def make_features(types):
    feats = []
    for name, count in types:
        for _ in range(count):
            if name == "bb":
                feats.append(dict(type="bb", **DEFAULT_HYPERS["bb"]))
            elif name == "sbb":        # a sparse boolean: rarely on, always observed
                feats.append(dict(type="bb", sparse=True, **DEFAULT_HYPERS["bb"]))
            elif name.startswith("dd"):
                dim = int(name[2:])
                feats.append(dict(type="dd", alphas=[DEFAULT_HYPERS["dd_alpha"]] * dim))
            elif name.startswith("dpd"):
                dim = int(name[3:])
                beta0 = 1.0 / (dim + 1)
                feats.append(dict(type="dpd", betas=[(1.0 - beta0) / dim] * dim, beta0=beta0,
                                  **DEFAULT_HYPERS["dpd"]))
            elif name == "gp":
                feats.append(dict(type="gp", **DEFAULT_HYPERS["gp"]))
            elif name == "nich":
                feats.append(dict(type="nich", **DEFAULT_HYPERS["nich"]))
            else:
                raise ValueError(name)
    rank = {"bb": 0, "dd": 1, "dpd": 2, "gp": 3, "nich": 4}
    feats.sort(key=lambda f: (rank[f["type"]], len(f.get("alphas", ()))))  # loom/schema.py:32-72
    return feats


def generate_kinds(feature_count, rng):
    """loom/generate.py:63-75: kind i gets 2**i features, then shuffled."""
    f2k = []
    for i in range(feature_count):
        f2k.extend([i] * (2 ** i))
        if len(f2k) >= feature_count:
            break
    f2k = np.asarray(f2k[:feature_count], dtype=np.uint32)
    rng.shuffle(f2k)
    return f2k


def sample_py_partition(n, alpha, d, rng):
    """Sequential Pitman-Yor seating (distributions Clustering::PitmanYor::
    sample_assignments, used at src/kind_kernel.cc:172-176)."""
    assign = np.empty(n, np.uint32)
    counts = np.zeros(64, np.float64)
    m = 0
    u = rng.random(n)
    for i in range(n):
        x = u[i] * (i + alpha)
        new = alpha + d * m
        if x < new or m == 0:
            g = m
            m += 1
            if m > len(counts):
                counts = np.concatenate([counts, np.zeros(len(counts))])
        else:
            c = np.cumsum(counts[:m] - d)
            g = min(int(np.searchsorted(c, x - new, side="right")), m - 1)
        counts[g] += 1
        assign[i] = g
    return assign, m


def sample_py_partition_fast(n, alpha, d, rng, sticks=4096):
    """The same Pitman-Yor partition law without the sequential seating loop (used for the 1M+ row configs, where
    9 kinds x 1M python iterations would dominate the bench set-up): draw the stick-breaking weights
    V_k ~ Beta(1 - d, alpha + (k + 1) d) of the PY process, then every row's table i.i.d. from them (the de Finetti
    form of the seating process; `sticks` leaves a negligible tail), and relabel tables by first appearance as
    the seating process does."""
    k = np.arange(1, sticks + 1, dtype=np.float64)
    v = rng.beta(1.0 - d, alpha + k * d)
    w = v * np.concatenate([[1.0], np.cumprod(1.0 - v[:-1])])
    w /= w.sum()
    raw = rng.choice(sticks, size=n, p=w)
    uniq, first = np.unique(raw, return_index=True)
    order = np.argsort(first)
    relabel = np.empty(sticks, np.int64)
    relabel[uniq[order]] = np.arange(len(uniq))
    return relabel[raw].astype(np.uint32), len(uniq)


def _sample_feature(f, assign, n_groups, rng):
    """values for every row from group-level latent parameters."""
    R = len(assign)
    t = f["type"]
    if t == "bb":
        p = rng.beta(0.2, 9.8, n_groups) if f.get("sparse") else rng.beta(f["alpha"], f["beta"], n_groups)
        return (rng.random(R) < p[assign]).astype(np.uint32)
    if t in ("dd", "dpd"):
        if t == "dd":
            conc = np.asarray(f["alphas"], np.float64)
        else:
            conc = f["alpha"] * np.asarray(f["betas"], np.float64)
        theta = rng.dirichlet(np.maximum(conc, 1e-3), n_groups)
        out = np.empty(R, np.uint32)
        order = np.argsort(assign, kind="stable")
        bounds = np.searchsorted(assign[order], np.arange(n_groups + 1))
        for g in range(n_groups):
            idx = order[bounds[g]:bounds[g + 1]]
            if len(idx):
                out[idx] = rng.choice(len(conc), size=len(idx), p=theta[g])
        return out
    if t == "gp":
        lam = rng.gamma(f["alpha"], 1.0 / f["inv_beta"], n_groups)
        return rng.poisson(lam[assign]).astype(np.uint32)
    if t == "nich":
        nu, s2 = f["nu"], f["sigmasq"]
        var = nu * s2 / rng.chisquare(nu, n_groups)
        mean = f["mu"] + np.sqrt(var / f["kappa"]) * rng.standard_normal(n_groups)
        x = mean[assign] + np.sqrt(var[assign]) * rng.standard_normal(R)
        return x.astype(np.float32).view(np.uint32)
    raise ValueError(t)


def generate(rows, types, density=0.5, seed=0, f2k=None, single_kind=False):
    """Returns (Model, RowBlock, truth) where truth["assign"][k] is the
    generating partition of kind k."""
    rng = np.random.default_rng(seed)
    feats = make_features(types)
    F = len(feats)
    if f2k is None:
        f2k = np.zeros(F, np.uint32) if single_kind else generate_kinds(F, rng)
    f2k = np.asarray(f2k, np.uint32)
    K = int(f2k.max()) + 1
    model = Model(feats, f2k, kind_py=[CLUSTERING] * K, topology=CLUSTERING)
    F8 = model.n8()
    cat8 = np.zeros((F8, rows), np.uint8)
    wide = np.zeros((F - F8, rows), np.uint32)
    truth = {"assign": [], "n_groups": []}
    for k in range(K):
        seat = sample_py_partition if rows <= 200_000 else sample_py_partition_fast
        assign, m = seat(rows, CLUSTERING[0], CLUSTERING[1], rng)
        truth["assign"].append(assign)
        truth["n_groups"].append(m)
        for f in model.kind_features(k):
            v = _sample_feature(feats[f], assign, m, rng)
            if f < F8:
                cat8[f] = v.astype(np.uint8)
            else:
                wide[f - F8] = v
    obs = rng.random((F, rows)) < density
    for f in range(F):
        if feats[f].get("sparse"):
            obs[f] = True
    return model, RowBlock(rows, cat8, wide, pack_mask(obs)), truth


def generate_config(name, rows=None):
    cfg = dict(CONFIGS[name])
    if rows is not None:
        cfg["rows"] = rows
    return generate(**cfg)
