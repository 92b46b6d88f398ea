"""Generator of tests/golden/reference_cat_kernel.json: multi-kind sweeps of the REFERENCE'S OWN CatKernel
(src/cat_kernel.hpp over the real CrossCat, ValueSplitter, Assignments, ProductModel, ProductMixture -- and Differ for
the tared case; compiled unmodified by `make -C oracle ref` into the git-ignored oracle/_ref/ref_cat_kernel) over the
functional mixtures of oracle/ref_stubs_mix/, the draw of every (row, kind) scripted from the oracle's own sweep: the
score vector of every draw, the final clustering counts, every row's group in the assignments queue (global id) and
mapped back through the id tracker (packed id), CrossCat::score_data.
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
import cat_pin_util as U                      # noqa: E402

exe = os.path.join(REPO, "oracle", "_ref", "ref_cat_kernel")
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
TYPE = {"bb": 0, "dd": 1, "dpd": 2, "gp": 3, "nich": 4}

cases = [
    dict(name="three_kinds", types=[("bb", 4), ("dd4", 2), ("dd256", 1), ("dpd5", 1), ("gp", 2), ("nich", 2)], f2k=[0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2],
         n_rows=60, density=0.7, kind_py=[2.0, 0.1], topology=[2.0, 0.2], passes=2, seed=71),
    dict(name="many_groups", types=[("bb", 3), ("nich", 3)], f2k=[0, 1, 0, 1, 0, 1], n_rows=50, density=1.0, kind_py=[6.0, 0.6], topology=[1.0, 0.0],
         passes=3, seed=72),
    dict(name="tared", types=[("bb", 6), ("dd4", 1), ("gp", 2), ("nich", 1)], f2k=[0, 0, 1, 1, 0, 1, 0, 1, 0, 1], n_rows=60, density=0.9,
         kind_py=[1.5, 0.2], topology=[2.0, 0.1], passes=2, seed=73, tare=True),
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
    mine = U.run(abi, case, model, table, tare)
    acts, picks = mine[0], mine[1]
    F = model.n_features
    text = "%r %r %r %r %d\n" % (case["topology"][0], case["topology"][1], case["kind_py"][0], case["kind_py"][1], F)
    for f, ft in enumerate(model.features):
        t = TYPE[ft["type"]]
        aux = ft.get("alphas") if t == 1 else ft.get("betas") if t == 2 else []
        hp = {0: [ft.get("alpha"), ft.get("beta"), 0, 0], 1: [0, 0, 0, 0], 2: [ft.get("gamma"), ft.get("alpha"), ft.get("beta0"), 0],
              3: [ft.get("alpha"), ft.get("inv_beta"), 0, 0], 4: [ft.get("mu"), ft.get("kappa"), ft.get("sigmasq"), ft.get("nu")]}[t]
        text += "%d %d %d %s %s\n" % (case["f2k"][f], t, len(aux), " ".join("%.9g" % np.float32(h) for h in hp),
                                      " ".join("%.9g" % np.float32(a) for a in aux))
    text += ("1\n" + " ".join(token(v) for v in tare_row) + "\n") if tare_row else "0\n"
    text += "%d\n" % len(table) + "\n".join(" ".join(token(v) for v in row) for row in table) + "\n"
    text += "%d\n" % len(acts) + "\n".join("%s %d" % ("R" if a & 1 else "A", a >> 1) for a in acts) + "\n"
    draws = [int(p) for i, a in enumerate(acts) if not a & 1 for p in picks[i]]
    text += "%d\n" % len(draws) + " ".join(str(p) for p in draws) + "\n"
    proc = subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE)
    assert proc.returncode == 0, (case["name"], proc.returncode)
    ref = json.loads(proc.stdout)
    n = U.compare(case, mine, ref)
    print(case["name"], "score vectors", n, "rows left", len(ref["rowids"]), "counts", [k["counts"] for k in ref["kinds"]], "score_data",
          ref["score_data"])
    return dict(case, tare_row=tare_row, reference=ref)


if __name__ == "__main__":
    out = [run_case(case) for case in cases]
    json.dump({"source": "/root/reference/src/cat_kernel.hpp over {cross_cat,product_model,product_mixture,product_value,differ}.cc compiled by "
                         "oracle/Makefile `ref` (g++ -O2) against oracle/ref_stubs_mix + ref_stubs_value + ref_stubs, linked to liboracle's om_* "
                         "per-feature functions, driven by oracle/ref_cat_kernel_main.cc", "cases": out},
              open(os.path.join(REPO, "tests", "golden", "reference_cat_kernel.json"), "w"), separators=(",", ":"))
