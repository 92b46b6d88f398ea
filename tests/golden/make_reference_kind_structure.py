"""Generator of tests/golden/reference_kind_structure.json: one whole kind-structure pass of the REFERENCE ITSELF --
its own CrossCat, KindKernel, KindProposer, ProductModel, ProductMixture, ValueSplitter and Assignments
(src/cross_cat.cc, kind_kernel.cc, kind_proposer.cc, product_model.cc, product_mixture.cc, product_value.cc and their
headers, compiled unmodified by `make -C oracle ref` into the git-ignored oracle/_ref/ref_kind_structure) over the
functional mixtures of oracle/ref_stubs_mix/ -- on seeded tables: rows go through KindKernel::add_row / remove_row,
try_run() closes every round, the kernel is destroyed at the end.  Every random decision is scripted from the oracle's
own run of the same pass (groups of the ADDs, uniforms of the block sampler, prior partitions of fresh ephemeral
kinds); the fixture holds what the reference COMPUTED: the score vector of every draw, L[f][k] of every round, the
feature -> kind map after every round, and at the end every kind's sufficient statistics and CrossCat::score_data.
Runs only where /root/reference exists (this container); the fixture it writes -- inputs and outputs -- is what travels."""
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
import structure_pin_util as U                # noqa: E402

exe = os.path.join(REPO, "oracle", "_ref", "ref_kind_structure")
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
abi.lib.lo_uniform.restype = ctypes.c_double
abi.lib.lo_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]
TYPE = {"bb": 0, "dd": 1, "dpd": 2, "gp": 3, "nich": 4}

cases = [
    dict(name="booleans", types=[("bb", 8)], f2k=[0, 0, 0, 1, 1, 1, 2, 2], n_rows=40, density=0.9, kind_py=[1.0, 0.1], topology=[1.0, 0.0],
         E=2, iterations=4, n_rounds=2, seed=51),
    dict(name="mixed", types=[("bb", 3), ("dd4", 2), ("dd256", 1), ("dpd5", 1), ("gp", 2), ("nich", 2)], f2k=[0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0],
         n_rows=50, density=0.7, kind_py=[2.0, 0.1], topology=[2.0, 0.2], E=3, iterations=8, n_rounds=2, seed=52),
    dict(name="one_kind", types=[("dd16", 3), ("gp", 2), ("nich", 2)], f2k=[0] * 7, n_rows=30, density=1.0, kind_py=[0.7, 0.3],
         topology=[3.0, 0.5], E=4, iterations=16, n_rounds=3, seed=53),
]


def token(v):
    return "-" if v is None else repr(v)


def run_case(case):
    model, table = U.build(case)
    mine = U.run(abi, case, model, table)
    F = model.n_features
    text = "%r %r %r %r %d %d %d\n" % (case["topology"][0], case["topology"][1], case["kind_py"][0], case["kind_py"][1], case["E"],
                                      case["iterations"], F)
    for f, ft in enumerate(model.features):
        t = TYPE[ft["type"]]
        aux = ft.get("alphas") if t == 1 else ft.get("betas") if t == 2 else []
        hp = {0: [ft.get("alpha"), ft.get("beta"), 0, 0], 1: [0, 0, 0, 0], 2: [ft.get("gamma"), ft.get("alpha"), ft.get("beta0"), 0],
              3: [ft.get("alpha"), ft.get("inv_beta"), 0, 0], 4: [ft.get("mu"), ft.get("kappa"), ft.get("sigmasq"), ft.get("nu")]}[t]
        text += "%d %d %d %s %s\n" % (case["f2k"][f], t, len(aux), " ".join("%.9g" % np.float32(h) for h in hp),
                                      " ".join("%.9g" % np.float32(a) for a in aux))
    text += "%d\n" % len(table) + "\n".join(" ".join(token(v) for v in row) for row in table) + "\n"
    text += "%d\n" % case["n_rounds"]
    for r, rd in enumerate(mine["rounds"]):
        acts = rd["actions"]
        text += "%d\n" % len(acts) + "\n".join("%s %d" % ("R" if a & 1 else "A", a >> 1) for a in acts) + "\n"
        draws = [int(p) for i, a in enumerate(acts) if not a & 1 for p in rd["picks"][i]]
        text += "%d\n" % len(draws) + " ".join(str(p) for p in draws) + "\n"
        n = case["iterations"] * F
        text += "%d\n" % n + " ".join("%.17g" % abi.lib.lo_uniform(case["seed"] + 100 + r, 1, i, 0) for i in range(n)) + "\n"
        text += "%d\n" % len(rd["seatings"]) + "".join("%d %s\n" % (len(s), " ".join(str(g) for g in s)) for s in rd["seatings"])
    proc = subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE)
    assert proc.returncode == 0, (case["name"], proc.returncode)
    ref = json.loads(proc.stdout)
    n = U.compare(case, model, mine, ref)
    print(case["name"], "score vectors", n, "kinds per round", [rd["kinds_scored"] for rd in ref["rounds"]],
          "maps", [rd["f2k"] for rd in ref["rounds"]], "final kinds", ref["final"]["n_kinds"], "score_data", ref["final"]["score_data"])
    return dict(case, reference=ref)


if __name__ == "__main__":
    out = [run_case(case) for case in cases]
    json.dump({"source": "/root/reference/src/{cross_cat,kind_kernel,kind_proposer,product_model,product_mixture,product_value}.cc compiled by "
                         "oracle/Makefile `ref` (g++ -O2) against oracle/ref_stubs_mix + ref_stubs_value + ref_stubs, linked to liboracle's om_* "
                         "per-feature functions, driven by oracle/ref_kind_structure_main.cc", "cases": out},
              open(os.path.join(REPO, "tests", "golden", "reference_kind_structure.json"), "w"), separators=(",", ":"))
