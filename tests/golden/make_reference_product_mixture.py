"""Generator of tests/golden/reference_product_mixture.json: the score vector of every ADD of a cat-kernel trajectory,
the final score_data and clustering counts, produced by the REFERENCE'S OWN ProductMixture_<true>
(src/product_mixture.cc + product_value.cc + differ.cc and the headers behind them, compiled unmodified by
`make -C oracle ref` into the git-ignored oracle/_ref/ref_product_mixture).  distributions' mixture classes are stood in
for by oracle/ref_stubs_mix/, which delegates the per-feature conjugate arithmetic to the oracle's own om_* functions:
the fixture therefore pins everything product_mixture.cc itself does -- prior + sum over observed features, group
births / deaths and the packed index every group gets, the tare cache, score_diff / add_diff / remove_diff on rows the
reference's own Differ compressed -- not the per-feature formulas (those are pinned to scipy, tests/test_oracle_math.py).
The trajectory (which packed group every ADD takes) is the oracle's own sweep of the same table.
Runs only where /root/reference exists (this container); the fixture it writes -- inputs and outputs -- is what travels."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "all", "ref"], stdout=subprocess.DEVNULL)
from loom_b200 import _abi, Engine            # noqa: E402
from loom_b200.engine import RowBlock         # noqa: E402
import mixture_pin_util as U                  # noqa: E402

exe = os.path.join(REPO, "oracle", "_ref", "ref_product_mixture")
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
TYPE = {"bb": 0, "dd": 1, "dpd": 2, "gp": 3, "nich": 4}

cases = [
    dict(name="booleans", types=[("bb", 6)], n_rows=60, density=0.8, alpha=1.0, d=0.0, passes=2, seed=41),
    dict(name="mixed", types=[("bb", 3), ("dd4", 2), ("dd256", 1), ("dpd5", 1), ("gp", 2), ("nich", 2)], n_rows=80, density=0.6,
         alpha=2.0, d=0.1, passes=2, seed=42),
    dict(name="sharp_py", types=[("dd16", 3), ("nich", 3)], n_rows=50, density=1.0, alpha=5.0, d=0.7, passes=3, seed=43),
    dict(name="tared", types=[("bb", 5), ("dd4", 1), ("gp", 2), ("nich", 1)], n_rows=70, density=0.9, alpha=1.5, d=0.2, passes=2,
         seed=44, tare=True),
]


def token(v):
    return "-" if v is None else repr(v)


def run_case(case):
    model, table = U.build(case)
    tare = tare_row = None
    if case.get("tare"):
        e = Engine(abi)
        e.set_model(model)
        e.set_rows(RowBlock.from_table(model, table))
        obs, val = e.make_tare()
        if obs.any():                             # (a random table may have no column with a majority value: run untared)
            tare = (obs, val)
            tare_row = [None if not obs[f] else int(val[f]) for f in range(model.n_features)]
    acts, picks, scores, score_data, counts = U.sweep(abi, case, model, table, tare)
    text = "%r %r %d\n" % (case["alpha"], case["d"], model.n_features)
    for ft in model.features:
        t = TYPE[ft["type"]]
        aux = ft.get("alphas") if t == 1 else ft.get("betas") if t == 2 else []
        hp = {0: [ft.get("alpha"), ft.get("beta"), 0, 0], 1: [0, 0, 0, 0], 2: [ft.get("gamma"), ft.get("alpha"), ft.get("beta0"), 0],
              3: [ft.get("alpha"), ft.get("inv_beta"), 0, 0], 4: [ft.get("mu"), ft.get("kappa"), ft.get("sigmasq"), ft.get("nu")]}[t]
        text += "%d %d %s %s\n" % (t, len(aux), " ".join("%.9g" % np.float32(h) for h in hp), " ".join("%.9g" % np.float32(a) for a in aux))
    text += ("1\n" + " ".join(token(v) for v in tare_row) + "\n") if tare_row else "0\n"
    text += "%d\n" % len(table) + "\n".join(" ".join(token(v) for v in row) for row in table) + "\n"
    text += "%d\n" % len(acts) + "\n".join("%s %d %d" % ("R" if a & 1 else "A", a >> 1, p) for a, p in zip(acts, picks)) + "\n"
    got = json.loads(subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout)
    assert len(got["adds"]) == len(scores)
    worst = 0.0
    for mine, ref in zip(scores, got["adds"]):
        assert len(mine) == len(ref), (case["name"], len(mine), len(ref))
        ref = np.asarray(ref)
        worst = max(worst, float(np.max(np.abs(mine - ref) / np.maximum(1.0, np.abs(ref)))))
    print(case["name"], "adds", len(scores), "max groups", max(len(s) for s in scores), "worst rel diff %.2e" % worst,
          "score_data", score_data, got["score_data"], "counts", list(counts), got["counts"])
    return dict(case, features=model.features, table=table, tare=tare_row, picks=[int(p) for p in picks],
                    adds=got["adds"], score_data=got["score_data"], counts=got["counts"], packed_to_global=got["packed_to_global"])


if __name__ == "__main__":
    out = [run_case(case) for case in cases]
    json.dump({"source": "/root/reference/src/product_mixture.cc + product_value.cc + differ.cc compiled by oracle/Makefile `ref` (g++ -O2) against "
                         "oracle/ref_stubs_mix + ref_stubs_value + ref_stubs, linked to liboracle's om_* per-feature functions, driven by "
                         "oracle/ref_product_mixture_main.cc", "cases": out},
              open(os.path.join(REPO, "tests", "golden", "reference_product_mixture.json"), "w"), separators=(",", ":"))
