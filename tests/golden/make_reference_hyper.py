"""Generator of tests/golden/reference_hyper.json: the hyperparameters chosen by the REFERENCE'S OWN HyperKernel::run
(src/hyper_kernel.cc over infer_grid.hpp / hyper_prior.hpp and the real CrossCat, ProductModel, ProductMixture; compiled
unmodified by `make -C oracle ref` into the git-ignored oracle/_ref/ref_hyper) over the functional mixtures of
oracle/ref_stubs_mix/ -- topology, every kind's clustering, every coordinate of every feature -- when its tasks draw
from the uniforms the oracle's tasks read.  Runs only where /root/reference exists (this container); the fixture it
writes -- inputs and outputs -- is what travels."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "all", "ref"], stdout=subprocess.DEVNULL)
from loom_b200 import _abi                    # noqa: E402
import hyper_pin_util as U                    # noqa: E402

exe = os.path.join(REPO, "oracle", "_ref", "ref_hyper")
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
abi.lib.lo_uniform.restype = ctypes.c_double
abi.lib.lo_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]
TYPE = {"bb": 0, "dd": 1, "dpd": 2, "gp": 3, "nich": 4}


def f32list(xs):
    return [float("%.9g" % np.float32(x)) for x in xs]


PY = [{"alpha": a, "d": d} for a in f32list([0.2, 0.7, 2.0, 6.0]) for d in f32list([0.0, 0.2, 0.5])]
GRIDS = {"topology": PY, "clustering": PY,
         "bb": {"alpha": f32list(np.logspace(-1, 1, 9)), "beta": f32list(np.logspace(-1, 1, 7))},
         "dd": {"alpha": f32list(np.logspace(-1, 1, 8))},
         "gp": {"alpha": f32list(np.logspace(-1, 2, 10)), "inv_beta": f32list(np.logspace(-2, 1, 11))},
         "nich": {"mu": f32list(np.linspace(-3, 3, 13)), "kappa": f32list(np.logspace(-1, 1, 6)), "sigmasq": f32list(np.logspace(-1, 1, 9)),
                  "nu": f32list(np.logspace(0, 1.5, 5))}}
ONE = dict(GRIDS, clustering=PY[:1], bb={"alpha": GRIDS["bb"]["alpha"], "beta": f32list([2.5])})    # single-point grids: set, no draw
cases = [
    dict(name="booleans", types=[("bb", 8)], f2k=[0, 0, 0, 1, 1, 1, 2, 2], n_rows=60, density=0.9, kind_py=[1.0, 0.1], topology=[1.0, 0.0],
         grids=GRIDS, seed=61),
    dict(name="mixed", types=[("bb", 3), ("dd4", 2), ("dd16", 1), ("gp", 3), ("nich", 3)], f2k=[0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1],
         n_rows=80, density=0.7, kind_py=[2.0, 0.1], topology=[2.0, 0.2], grids=GRIDS, seed=62),
    dict(name="single_points", types=[("bb", 4), ("gp", 2), ("nich", 2)], f2k=[0] * 8, n_rows=40, density=1.0, kind_py=[0.7, 0.3],
         topology=[3.0, 0.5], grids=ONE, seed=63),
]


def token(v):
    return "-" if v is None else repr(v)


def grid_text(values):
    return "%d %s\n" % (len(values), " ".join("%.9g" % v for v in values))


def run_case(case):
    model, table = U.build(case)
    mine = U.run(abi, case, model, table)
    picks = mine[0]
    F, K = model.n_features, model.n_kinds
    text = "%r %r %r %r %d\n" % (case["topology"][0], case["topology"][1], case["kind_py"][0], case["kind_py"][1], F)
    for f, ft in enumerate(model.features):
        t = TYPE[ft["type"]]
        aux = ft.get("alphas") if t == 1 else ft.get("betas") if t == 2 else []
        hp = {0: [ft.get("alpha"), ft.get("beta"), 0, 0], 1: [0, 0, 0, 0], 2: [ft.get("gamma"), ft.get("alpha"), ft.get("beta0"), 0],
              3: [ft.get("alpha"), ft.get("inv_beta"), 0, 0], 4: [ft.get("mu"), ft.get("kappa"), ft.get("sigmasq"), ft.get("nu")]}[t]
        text += "%d %d %d %s %s\n" % (case["f2k"][f], t, len(aux), " ".join("%.9g" % np.float32(h) for h in hp),
                                      " ".join("%.9g" % np.float32(a) for a in aux))
    text += "%d\n" % len(table)
    for r, row in enumerate(table):
        text += " ".join(token(v) for v in row) + " " + " ".join(str(int(p)) for p in picks[r]) + "\n"
    g = case["grids"]
    for name in ("topology", "clustering"):
        text += "%d %s\n" % (len(g[name]), " ".join("%.9g %.9g" % (p["alpha"], p["d"]) for p in g[name]))
    for t, params in (("bb", ("alpha", "beta")), ("dd", ("alpha",)), ("gp", ("alpha", "inv_beta")), ("nich", ("mu", "kappa", "sigmasq", "nu"))):
        for p in params:
            text += grid_text(g[t][p])
    for task in range(1 + K + F):
        text += "300 " + " ".join("%.17g" % abi.lib.lo_uniform(case["seed"] + 7, 2, task, j) for j in range(300)) + "\n"
    proc = subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE)
    assert proc.returncode == 0, (case["name"], proc.returncode)
    ref = json.loads(proc.stdout)
    U.compare(case, mine, ref)
    print(case["name"], "draws", ref["draws"], "topology", ref["topology"], "kinds", ref["kinds"], "first features", ref["features"][:2])
    return dict(case, reference=ref)


if __name__ == "__main__":
    out = [run_case(case) for case in cases]
    json.dump({"source": "/root/reference/src/{hyper_kernel,cross_cat,product_model,product_mixture,product_value}.cc compiled by oracle/Makefile "
                         "`ref` (g++ -O2) against oracle/ref_stubs_mix + ref_stubs_value + ref_stubs, linked to liboracle's om_* per-feature "
                         "functions, driven by oracle/ref_hyper_main.cc; uniforms of task t = lo_uniform(seed + 7, 2, t, draw)", "cases": out},
              open(os.path.join(REPO, "tests", "golden", "reference_hyper.json"), "w"), separators=(",", ":"))
