"""Generator of tests/golden/reference_schedules.json: ADD / REMOVE decision sequences and batch boundaries produced by
the REFERENCE'S OWN schedule classes (src/schedules.hpp, compiled in place by `make -C oracle ref` into the git-ignored
oracle/_ref/ref_schedules) driven through the loop of Loom::infer_kind_structure_sequential (src/loom.cc:217-262).
Runs only where /root/reference exists (this container); the fixture it writes is what travels."""
import json
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "ref"], stdout=subprocess.DEVNULL)
exe = os.path.join(REPO, "oracle", "_ref", "ref_schedules")
cases = [
    # (extra_passes, small_data_size, big_data_size, rows[, max_actions])
    (0, 1, 1, 30), (1, 1e9, 1e9, 50), (2, 10, 100, 40), (0.5, 5, 500, 120), (10, 4, 64, 100), (3.7, 0, 1000, 200),
    (500, 4e3, 1e9, 12, 9000),                        # loom's defaults (loom/config.py): the head of a run (truncated)
    (20, 8, 600, 500, 30000),                         # an accelerating schedule crossing both thresholds (truncated)
    (1, 100, 100, 400), (60, 2, 4, 6),
]
out = []
for c in cases:
    args = [repr(float(c[0])), repr(float(c[1])), repr(float(c[2])), str(c[3])] + ([str(c[4])] if len(c) > 4 else [])
    line = subprocess.check_output([exe] + args).decode()
    d = json.loads(line)
    d.update(extra_passes=float(c[0]), small=float(c[1]), big=float(c[2]), max_actions=c[4] if len(c) > 4 else None)
    out.append(d)
json.dump({"source": "/root/reference/src/schedules.hpp compiled by oracle/Makefile `ref` (g++ -O2), driven by oracle/ref_schedules_main.cc",
           "cases": out}, open(os.path.join(REPO, "tests", "golden", "reference_schedules.json"), "w"), indent=0)
print("%d cases, %d decisions" % (len(out), sum(len(d["decisions"]) for d in out)))
