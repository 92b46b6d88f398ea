"""Generator of tests/golden/reference_differ.json: the tare row and every sparsified row produced by the REFERENCE'S
OWN loom::Differ (src/differ.cc + src/product_value.cc, compiled unmodified by `make -C oracle ref` into the git-ignored
oracle/_ref/ref_differ; protoc's classes, libprotobuf's streams and the un-vendored distributions headers are stood in
for by oracle/ref_stubs_value/) on seeded tables: Differ::add_rows (what `loom tare` runs) and Differ::compress_rows
(`loom sparsify`), including the representation ValueSchema::normalize_small picks for each written value.
Runs only where /root/reference exists (this container); the fixture it writes -- inputs and outputs -- is what travels."""
import json
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "ref"], stdout=subprocess.DEVNULL)
exe = os.path.join(REPO, "oracle", "_ref", "ref_differ")


def f32(x):
    return float("%.9g" % np.float32(x))


def mixed(rng, R):
    """booleans 6, counts = dd4 x 2, dd256, dpd5 x 2, gp x 3, reals 2: the shape of tests/test_tare.py's differ_case"""
    types = [("bb", 6), ("dd4", 2), ("dd256", 1), ("dpd5", 2), ("gp", 3), ("nich", 2)]
    cols = []
    for j in range(4):
        cols.append((rng.random(R) < 0.04).astype(int))          # sparse booleans: tare 0
    cols.append((rng.random(R) < 0.97).astype(int))              # nearly always true: tare 1
    cols.append((rng.random(R) < 0.5).astype(int))               # no majority among the observed + unobserved rows
    cols.append(np.where(rng.random(R) < 0.8, 3, rng.integers(0, 4, R)))
    cols.append(rng.integers(0, 4, R))
    cols.append(np.where(rng.random(R) < 0.8, 200, 3))           # the mode 200 is outside [0, 16): never tared
    cols.append(np.where(rng.random(R) < 0.9, 2, rng.integers(0, 5, R)))
    cols.append(rng.integers(0, 5, R))
    cols.append(np.where(rng.random(R) < 0.7, 5, rng.poisson(3.0, R)))
    cols.append(np.where(rng.random(R) < 0.7, 16, rng.poisson(3.0, R)))   # 16: the first value the summary ignores
    cols.append(np.where(rng.random(R) < 0.55, 0, rng.poisson(8.0, R)))
    cols.append(rng.standard_normal(R))
    cols.append(rng.standard_normal(R) * 100)
    obs = rng.random((len(cols), R)) < 0.85
    obs[:5] |= rng.random((5, R)) < 0.9
    table = []
    for r in range(R):
        row = []
        for f, c in enumerate(cols):
            row.append(None if not obs[f, r] else (f32(c[r]) if f >= 14 else int(c[r])))
        table.append(row)
    return types, table


def wide_booleans(rng, R, F):
    """a wide, sparse boolean table: rows whose pos holds under 10 % of the columns are written SPARSE, empty ones NONE"""
    ones = rng.random((R, F)) < 0.03
    obs = rng.random((R, F)) < 0.98
    obs[0] = True
    ones[0] = False                                              # a row equal to the tare: pos NONE, neg NONE
    obs[1] = False                                               # a row observing nothing: neg = the whole tare
    return [("bb", F)], [[None if not obs[r, f] else int(ones[r, f]) for f in range(F)] for r in range(R)]


def edges():
    """ties go to the lower value; exactly half of the rows is not "more than half" (src/differ.hpp:59, differ.cc:197-199)"""
    return [("bb", 3), ("gp", 2)], [[1, 1, 1, 0, 7], [0, 1, 0, 0, 7], [1, None, 0, 1, 7], [0, 1, None, 1, None]]


def no_tare(rng, R):
    """no column has a majority value: the rows pass through, only normalised (src/differ.cc:183-188)"""
    table = [[int(rng.integers(0, 2)) if rng.random() < 0.5 else None, int(rng.integers(0, 16)), f32(rng.standard_normal())]
             for _ in range(R)]
    table[3] = [1, 2, 0.5]                                       # ALL
    table[4] = [None, None, None]                                # NONE
    return [("bb", 1), ("gp", 1), ("nich", 1)], table


def run_case(name, types, table, verbose=True):
    sizes = [0, 0, 0]
    for t, n in types:
        sizes[0 if t == "bb" else 2 if t == "nich" else 1] += n
    text = "%d %d %d %d\n" % (sizes[0], sizes[1], sizes[2], len(table))
    text += "\n".join(" ".join("-" if v is None else repr(v) for v in row) for row in table) + "\n"
    got = json.loads(subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout)
    assert len(got["rows"]) == len(table)
    if verbose:
        print(name, "tare over", len(got["tare"]["features"]), "of", sum(sizes), "columns;",
              sorted(set(r["pos"]["sparsity"] for r in got["rows"])), sorted(set(r["neg"]["sparsity"] for r in got["rows"])))
    return dict(name=name, types=types, table=table, tare=got["tare"], rows=got["rows"])


if __name__ == "__main__":
    rng = np.random.default_rng(20)
    cases = [("mixed", ) + mixed(rng, 400), ("wide_booleans", ) + wide_booleans(rng, 60, 120), ("edges", ) + edges(),
             ("no_tare", ) + no_tare(rng, 12)]
    out = [run_case(*case) for case in cases]
    json.dump({"source": "/root/reference/src/differ.cc + src/product_value.cc compiled by oracle/Makefile `ref` (g++ -O2) against "
                         "oracle/ref_stubs_value, driven by oracle/ref_differ_main.cc",
               "cases": out}, open(os.path.join(REPO, "tests", "golden", "reference_differ.json"), "w"), separators=(",", ":"))
