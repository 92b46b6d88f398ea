"""Shared by tests/golden/make_reference_product_mixture.py (the generator) and tests/test_reference_pins.py: a
single-kind engine swept by the cat kernel with its debug taps on -- the trajectory (packed group of every ADD and
REMOVE) and the score vector of every ADD that the compiled reference ProductMixture was driven with / must agree
with (oracle/ref_product_mixture_main.cc).  Test infrastructure."""
import os

import numpy as np

from loom_b200 import Engine, synth, sweep_actions
from loom_b200.engine import Model, RowBlock

STRIDE = int(os.environ.get("PIN_STRIDE", 64))     # score slots per ADD in the debug buffer; the committed cases stay below 64,
                                                  # the fuzzer's longer trajectories ask for more


def build(case):
    """-> (model, table): varied hypers per feature (the fixture stores them), a seeded table with the case's density"""
    feats = synth.make_features([tuple(t) for t in case["types"]])
    rng = np.random.RandomState(case["seed"])
    for f in feats:                       # hypers away from the defaults, so that a swapped pair of them would show
        if f["type"] == "bb":
            f.update(alpha=float(np.float32(rng.uniform(0.3, 3))), beta=float(np.float32(rng.uniform(0.3, 3))))
        elif f["type"] == "dd":
            f["alphas"] = [float(np.float32(a)) for a in rng.uniform(0.2, 2.0, len(f["alphas"]))]
        elif f["type"] == "gp":
            f.update(alpha=float(np.float32(rng.uniform(0.5, 4))), inv_beta=float(np.float32(rng.uniform(0.2, 2))))
        elif f["type"] == "nich":
            f.update(mu=float(np.float32(rng.normal())), kappa=float(np.float32(rng.uniform(0.3, 3))),
                     sigmasq=float(np.float32(rng.uniform(0.5, 3))), nu=float(np.float32(rng.uniform(1, 5))))
    model = Model(feats, np.zeros(len(feats), np.uint32), kind_py=[(case["alpha"], case["d"])], topology=(case["alpha"], case["d"]))
    _, rows, _ = synth.generate(case["n_rows"], [tuple(t) for t in case["types"]], density=case["density"], seed=case["seed"],
                                single_kind=True)
    obs = rows.observed()
    F8 = model.n8()
    table = []
    for r in range(case["n_rows"]):
        row = []
        for f, ft in enumerate(feats):
            if not obs[f, r]:
                row.append(None)
            elif f < F8:
                row.append(int(rows.cat8[f, r]))
            elif ft["type"] == "nich":
                row.append(float("%.9g" % rows.wide[f - F8, r:r + 1].view(np.float32)[0]))
            else:
                row.append(int(rows.wide[f - F8, r]))
        table.append(row)
    if case.get("tare"):                  # make the first booleans nearly constant so that a tare row exists
        for r in range(case["n_rows"]):
            for f in range(min(4, F8)):
                if feats[f]["type"] == "bb" and rng.uniform() < 0.9:
                    table[r][f] = 1 if f == 1 else 0
    return model, table


def actions_of(case):
    """ADD every row, then REMOVE + ADD passes (src/loom.cc:502-505), then REMOVE the first half: births and deaths"""
    R = case["n_rows"]
    acts = [sweep_actions(R)]
    for _ in range(case["passes"]):
        acts.append(sweep_actions(R, add=True, remove=True))
    acts.append(sweep_actions(R // 2, add=False, remove=True))
    return np.concatenate(acts)


def sweep(abi, case, model, table, tare=None):
    """-> (actions, picks [n], scores: one float32 vector per ADD, score_data, counts)"""
    e = Engine(abi)
    e.set_model(model)
    block = RowBlock.from_table(model, table)
    e.set_rows(block)
    if tare is not None:                  # the same rows as a tare-compressed block (lb_set_rows_diff)
        obs, val = tare
        d = e.sparsify(obs, val)
        d["row_has_tare"] = np.ones(len(table), np.uint8)
        e.set_rows_diff(**d)
    acts = actions_of(case)
    picks, scores = e.cat_sweep(acts, seed=case["seed"], debug=True, debug_stride=STRIDE)
    out = []
    for i, a in enumerate(acts):
        if not a & 1:
            s = scores[i, 0]
            out.append(s[~np.isnan(s)])
    info = e.kind_info(0)
    counts, _, _ = e.get_groups(0)
    return acts, picks[:, 0], out, float(e.score_data()), np.asarray(counts)[:info.n_groups]
