"""Generator of tests/golden/reference_kind_proposer.json: every draw of the block Pitman-Yor sampler and the final
feature -> kind assignment produced by the REFERENCE'S OWN KindProposer::infer_assignments (src/kind_proposer.cc,
compiled unmodified by `make -C oracle ref` into the git-ignored oracle/_ref/ref_kind_proposer; the un-vendored
distributions library and the mixture that produces the scores are stood in for by oracle/ref_stubs/) on seeded score
matrices.  The uniforms fed to the reference are the ones the oracle's sampler will read from its Philox stream
(lo_uniform(seed, 1, iteration * F + f, 0)), so the two can be compared draw for draw.
Runs only where /root/reference exists (this container); the fixture it writes is what travels."""
import ctypes
import json
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "all", "ref"], stdout=subprocess.DEVNULL)
exe = os.path.join(REPO, "oracle", "_ref", "ref_kind_proposer")
oracle = ctypes.CDLL(os.path.join(REPO, "oracle", "libloom_oracle.so"))
oracle.lo_uniform.restype = ctypes.c_double
oracle.lo_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]

cases = [
    # (alpha, d, F, K, iterations, score spread, seed, initial assignment)
    (1.0, 0.0, 6, 3, 4, 1.0, 11, "one"),          # everything in kind 0, two empty kinds
    (0.5, 0.25, 12, 5, 8, 3.0, 12, "cycle"),
    (2.0, 0.5, 40, 12, 32, 10.0, 13, "random"),   # C2-like: ~8 kinds + ephemeral ones, loom's 32 iterations
    (4.0, 0.9, 64, 41, 32, 30.0, 14, "few"),      # C3's K + E = 41, sharp likelihoods
    (0.1, 0.0, 20, 20, 16, 0.1, 15, "identity"),  # no empty kind at the start, flat likelihoods
    (1.5, 0.3, 256, 41, 32, 100.0, 16, "random"), # C3's feature count
    (1.0, 0.1, 9, 2, 10, 700.0, 17, "one"),       # likelihoods that underflow float32 next to the maximum
    (3.0, 0.6, 30, 7, 5, 5.0, 18, "cycle"),
]
out = []
for alpha, d, F, K, iterations, spread, seed, init in cases:
    rng = np.random.RandomState(seed)
    scores = (rng.standard_normal((F, K)) * spread - 1000.0).astype(np.float32)    # marginal log-likelihoods are large negatives
    if init == "one":
        assign = np.zeros(F, np.uint32)
    elif init == "cycle":
        assign = (np.arange(F) % K).astype(np.uint32)
    elif init == "identity":
        assign = np.arange(F, dtype=np.uint32) % K
    elif init == "few":
        assign = rng.randint(0, 9, F).astype(np.uint32)
    else:
        assign = rng.randint(0, K, F).astype(np.uint32)
    uniforms = [oracle.lo_uniform(seed, 1, i, 0) for i in range(iterations * F)]
    text = "%r %r %d %d %d\n" % (alpha, d, F, K, iterations)
    text += " ".join(str(int(a)) for a in assign) + "\n"
    text += "\n".join(" ".join("%.9g" % v for v in row) for row in scores) + "\n"
    text += "%d\n" % len(uniforms) + " ".join("%.17g" % u for u in uniforms) + "\n"
    got = json.loads(subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout)
    assert len(got["picks"]) == iterations * F
    out.append(dict(alpha=alpha, d=d, F=F, K=K, iterations=iterations, seed=seed, initial=[int(a) for a in assign],
                    scores=[[float("%.9g" % v) for v in row] for row in scores],
                    picks=got["picks"], totals=got["totals"], assignments=got["assignments"]))
json.dump({"source": "/root/reference/src/kind_proposer.cc compiled by oracle/Makefile `ref` (g++ -O2) against oracle/ref_stubs, "
                     "driven by oracle/ref_kind_proposer_main.cc; uniforms = lo_uniform(seed, 1, draw, 0)",
           "cases": out}, open(os.path.join(REPO, "tests", "golden", "reference_kind_proposer.json"), "w"), separators=(",", ":"))
print("%d cases, %d draws" % (len(out), sum(len(c["picks"]) for c in out)))
