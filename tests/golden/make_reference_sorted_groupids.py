"""Generator of tests/golden/reference_sorted_groupids.json: the dump order of groups -- CrossCat::get_sorted_groupids
(src/cross_cat.cc:212-236, compiled unmodified by `make -C oracle ref` into the git-ignored oracle/_ref/ref_sorted_groupids)
-- on seeded count vectors with many tied groups, below and above the 16 elements at which libstdc++'s std::sort stops
being an insertion sort.  Runs only where /root/reference exists (this container); the fixture is what travels."""
import json
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "all", "ref"], stdout=subprocess.DEVNULL)
exe = os.path.join(REPO, "oracle", "_ref", "ref_sorted_groupids")
rng = np.random.RandomState(81)
kinds = [[3, 0, 5, 5, 1, 1, 1, 0, 2], [1] * 16 + [0], [1] * 17 + [0], [2, 1] * 20 + [0], [0], [7]]
for n in (5, 16, 17, 33, 64, 200, 1000):
    kinds.append(rng.randint(0, 4, n).tolist())                       # counts 0..3: ties everywhere, empties inside
    kinds.append((rng.zipf(1.6, n) % 50).tolist())                    # a Pitman-Yor-like profile: a few large, many singletons
text = "%d\n" % len(kinds) + "".join("%d %s\n" % (len(c), " ".join(str(x) for x in c)) for c in kinds)
orders = json.loads(subprocess.run([exe], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout)
unstable = sum(o != sorted([g for g in range(len(c)) if c[g]], key=lambda g: -c[g]) for c, o in zip(kinds, orders))
print(len(kinds), "kinds,", unstable, "of them ordered differently from a stable sort")
json.dump({"source": "/root/reference/src/cross_cat.cc compiled by oracle/Makefile `ref` (g++ -O2, libstdc++ of GCC 13), driven by "
                     "oracle/ref_sorted_groupids_main.cc", "kinds": [{"counts": c, "order": o} for c, o in zip(kinds, orders)]},
          open(os.path.join(REPO, "tests", "golden", "reference_sorted_groupids.json"), "w"), separators=(",", ":"))
