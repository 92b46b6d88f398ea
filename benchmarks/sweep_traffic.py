"""ncu CSV of every sweep-kernel launch of a bench run -> profiles/r2_k_cat_sweep_traffic.json (the LAST complete
REMOVE+ADD sweep: one launch of each class with its full grid).
usage: sweep_traffic.py CSV OUT.json WORKLOAD ROWS CHAINS KINDS"""
import collections
import csv
import json
import sys

src, out, workload, R, S, K = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
h = rows[0]
launches = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    e = launches.setdefault(int(d["ID"]), {"name": d["Kernel Name"].split("(")[0], "grid": int(d["Grid Size"].strip("()").split(",")[0])})
    e[d["Metric Name"]] = float(d["Metric Value"])
ids = sorted(launches)
full = {}
for i in ids:                        # the largest grid seen per kernel = the full class
    n = launches[i]["name"]
    full[n] = max(full.get(n, 0), launches[i]["grid"])
last = {}
for i in ids:
    l = launches[i]
    if l["grid"] == full[l["name"]]:
        last[l["name"]] = l
per = {n: {"dram_bytes": l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"], "ms_alone_under_ncu": l["gpu__time_duration.sum"] / 1e6,
           "grid": l["grid"]} for n, l in last.items()}
tot = sum(v["dram_bytes"] for v in per.values())
history = [{"id": i, "kernel": launches[i]["name"], "grid": launches[i]["grid"], "ms": launches[i]["gpu__time_duration.sum"] / 1e6,
            "dram_GB": (launches[i]["dram__bytes_read.sum"] + launches[i]["dram__bytes_write.sum"]) / 1e9} for i in ids]
json.dump({"workload": workload, "rows": R, "chains": S, "kinds_per_chain": K, "row_actions_per_sweep_per_chain": 2 * R,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum of the four class kernels of the LAST complete REMOVE+ADD sweep of "
                   "`python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline` (ncu --metrics only; kernels serialised by ncu)",
           "per_kernel": per, "dram_bytes_per_launch": tot, "dram_bytes_per_chain_row_action": tot / (S * 2.0 * R),
           "every_sweep_launch_of_the_run": history}, open(out, "w"), indent=1)
print("dram bytes per sweep %.4g = %.0f B per chain row action" % (tot, tot / (S * 2.0 * R)))
