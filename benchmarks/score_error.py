"""Per-feature-type score error of the device against the float64 oracle (parity check 2): for each type a
single-kind model of 12 features is swept once on the device, the oracle is given the device's groups, and both score
every row against every group (K1).  Reports max and 99.9th-percentile of |device - oracle| / max(1, |oracle|) per type,
and the error of the float moments the device keeps (GP log_prod, NICH mean / count_times_variance) against a float64
recomputation from the rows.  usage: score_error.py ROWS   -> one JSON line"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions


def measure(cuda, oracle, rows_n, types=("bb", "dd16", "dd256", "dpd8", "gp", "nich")):
    out = {}
    for t in types:
        model, rows, _ = synth.generate(rows_n, [(t, 12)], density=0.7, seed=17, single_kind=True)
        g = Engine(cuda); g.set_model(model); g.set_rows(rows)
        g.cat_sweep(sweep_actions(rows_n), seed=3)
        counts, ints, flts = g.get_groups(0)
        keep = counts > 0
        o = Engine(oracle); o.set_model(model); o.set_rows(rows)
        o.set_groups(0, int(keep.sum()), counts[keep], ints[:, keep], flts[:, keep])
        g2 = Engine(cuda); g2.set_model(model); g2.set_rows(rows)      # same packing as the oracle's copy
        g2.set_groups(0, int(keep.sum()), counts[keep], ints[:, keep], flts[:, keep])
        a, b = g2.score_rows(0).astype(np.float64), o.score_rows(0).astype(np.float64)
        err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
        entry = {"groups": int(keep.sum()), "scores": int(err.size), "max_err": float(err.max()),
                 "p999_err": float(np.quantile(err, 0.999))}
        if t in ("gp", "nich"):                                            # float moments against a recomputation
            o.set_groups(0, int(keep.sum()))
            o.set_assignments(0, remap(g.get_assignments(0), keep))
            o.accumulate_rows()
            _, _, f2 = o.get_groups(0)
            n = int(keep.sum())
            rel = np.abs(flts[:, keep] - f2[:, :n]) / np.maximum(1.0, np.abs(f2[:, :n]))
            entry["moment_max_rel_err"] = float(rel.max())
        out[t] = entry
    return out


def remap(assign, keep):
    """packed ids of the device's sweep -> ids after the empty groups are dropped"""
    new = np.full(len(keep), _abi.UNASSIGNED, np.uint32)
    new[keep] = np.arange(int(keep.sum()), dtype=np.uint32)
    out = assign.copy()
    ok = assign != _abi.UNASSIGNED
    out[ok] = new[assign[ok]]
    return out


if __name__ == "__main__":
    cuda = _abi.load_cuda()
    oracle = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
    res = measure(cuda, oracle, int(sys.argv[1]))
    res["tolerance"] = "1e-5 * max(1, |oracle score|) (north star); moments: float64 on both sides"
    print(json.dumps(res))
