"""Generator of tests/golden/reference_kind_kernel.json: the feature -> kind map and the kind count after every step of
a KindKernel's life -- construction, try_run() calls, destruction -- produced by the REFERENCE'S OWN
src/kind_kernel.{hpp,cc} + src/kind_proposer.{hpp,cc} (compiled unmodified by `make -C oracle ref` into the git-ignored
oracle/_ref/ref_kind_kernel; CrossCat, the mixtures and the un-vendored distributions library are stood in for by
oracle/ref_stubs_kind/, ref_stubs_value/, ref_stubs/) on seeded score matrices.  The uniforms fed to the reference are
the ones the oracle's block sampler reads from its Philox stream in the same round, so the two make the same moves and
what is compared is the bookkeeping after them: which kinds die, how swap-removal renumbers the survivors, where the
new ephemeral kinds go.
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
import kind_pin_util                          # noqa: E402

exe = os.path.join(REPO, "oracle", "_ref", "ref_kind_kernel")
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
abi.lib.lo_uniform.restype = ctypes.c_double
abi.lib.lo_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]

cases = [
    # types, initial feature -> kind map, E, iterations, rounds, spread of the scores, (alpha, d), rows, seed
    dict(types=[("bb", 4), ("gp", 2)], f2k=[0, 0, 1, 1, 2, 2], E=2, iterations=4, n_rounds=3, spread=4.0, alpha=1.0, d=0.1, n_rows=40, seed=31),
    dict(types=[("bb", 6), ("dd4", 3), ("nich", 3)], f2k=[0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3], E=4, iterations=8, n_rounds=4, spread=8.0, alpha=2.0, d=0.3, n_rows=50, seed=32),
    dict(types=[("bb", 20)], f2k=[f % 7 for f in range(20)], E=8, iterations=32, n_rounds=5, spread=20.0, alpha=0.5, d=0.0, n_rows=30, seed=33),
    dict(types=[("dd16", 5), ("gp", 5)], f2k=[0] * 10, E=3, iterations=2, n_rounds=3, spread=1.0, alpha=4.0, d=0.6, n_rows=30, seed=34),
    dict(types=[("bb", 16)], f2k=list(range(16)), E=1, iterations=6, n_rounds=4, spread=6.0, alpha=1.0, d=0.5, n_rows=20, seed=35),
]
out = []
for case in cases:
    rng = np.random.RandomState(case["seed"])
    F = len(case["f2k"])
    rounds = []

    def draw(r, n_kinds):
        S = (rng.standard_normal((n_kinds, F)) * case["spread"] - 50.0).astype(np.float32)
        rounds.append([[float("%.9g" % v) for v in row] for row in S])
        return rounds[-1]

    states = kind_pin_util.life_cycle(abi, case, draw)
    text = "%r %r %d %d %d %d %d\n" % (case["alpha"], case["d"], F, max(case["f2k"]) + 1, case["E"], case["iterations"], case["n_rows"])
    text += " ".join(str(k) for k in case["f2k"]) + "\n%d\n" % case["n_rounds"]
    for r, S in enumerate(rounds):
        n = case["iterations"] * F
        uniforms = [abi.lib.lo_uniform(case["seed"] + 100 + r, 1, i, 0) for i in range(n)]
        text += "%d\n" % len(S) + "\n".join(" ".join("%.9g" % v for v in row) for row in S) + "\n"
        text += "%d\n" % n + " ".join("%.17g" % u for u in uniforms) + "\n"
    got = json.loads(subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout)
    steps = got["steps"]
    assert len(steps) == len(states)
    print(case["seed"], [(s["n_kinds"], s.get("change_count"), s.get("birth_count"), s.get("death_count")) for s in steps])
    for mine, ref in zip(states, steps):
        assert mine[0] == ref["step"]
        if (mine[1], mine[2]) != (ref["f2k"], ref["n_kinds"]):
            print("  MISMATCH at", ref["step"], "\n   oracle   ", mine[1], mine[2], "\n   reference", ref["f2k"], ref["n_kinds"])
    out.append(dict(case, scores=rounds, steps=steps))
json.dump({"source": "/root/reference/src/kind_kernel.cc + kind_proposer.cc compiled by oracle/Makefile `ref` (g++ -O2) against "
                     "oracle/ref_stubs_kind + ref_stubs_value + ref_stubs, driven by oracle/ref_kind_kernel_main.cc; "
                     "uniforms of round r = lo_uniform(seed + 100 + r, 1, draw, 0)",
           "cases": out}, open(os.path.join(REPO, "tests", "golden", "reference_kind_kernel.json"), "w"), separators=(",", ":"))
