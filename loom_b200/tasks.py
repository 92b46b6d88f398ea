"""`loom.tasks.infer` on the device engine: several samples of one dataset, from loom's files to
loom's files (loom/tasks.py:141-245, loom/generate.py:63-135, loom/runner.py:200-243).

The reference starts one `loom_infer` process per sample (`parallel_map(_infer_one, ...)`), each
with its own seed, its own initial model and its own shuffle of the rows.  Here the samples are
chains on ONE device: the rows file is decoded once into a shared columnar block, every chain
gets its own initial kind structure / hypers and its own row order, and each batch of the
schedule is one launch over all chains (`lb_cat_sweep_streams`).  Per sample the outputs are the
files `loom_infer` writes: model.pb.gz, groups/mixture.N.pbs.gz, assign.pbs.gz.

For a single sample `loom_b200/host/loom_infer` (C++) is the drop-in for the process itself.
"""
import os

import numpy as np

from . import _abi, pbio
from .engine import Engine, Model
from .infer import infer_multi_pass, infer_single_pass
from .synth import CLUSTERING, generate_kinds


def sample_grid(grid, rng):
    """loom/generate.py:50-60"""
    if isinstance(grid, dict):
        return {k: sample_grid(v, rng) for k, v in grid.items()}
    return grid[int(rng.integers(len(grid)))]


def generate_init(features, hyper_prior, seed):
    """loom.generate.generate_init (loom/generate.py:99-135, import_features :92-117): an exponential
    kind structure, hypers drawn from the grids, PY(2, 0.1) clusterings.  `features` gives each
    feature's type and size: {"type": "bb"|"gp"|"nich"} | {"type": "dd", "dim": n} |
    {"type": "dpd", "betas": [...], "beta0": b} (a DPD feature keeps the dictionary it was given)."""
    rng = np.random.default_rng(seed)
    feats = []
    for f in features:
        t = f["type"]
        if t == "dd":
            dim = f.get("dim") or len(f["alphas"])
            feats.append({"type": "dd", "alphas": [float(sample_grid(hyper_prior["dd"]["alpha"], rng)) for _ in range(dim)]})
        elif t == "dpd":
            raw = sample_grid(hyper_prior["dpd"], rng)
            feats.append({"type": "dpd", "gamma": float(raw["gamma"]), "alpha": float(raw["alpha"]),
                          "betas": list(f["betas"]), "beta0": f["beta0"]})
        else:
            feats.append(dict({"type": t}, **{k: float(v) for k, v in sample_grid(hyper_prior[t], rng).items()}))
    f2k = generate_kinds(len(feats), rng)
    n_kinds = int(f2k.max()) + 1
    return Model(feats, f2k, kind_py=[CLUSTERING] * n_kinds, topology=CLUSTERING)


def _dump_sample(engine, model_file, rowids, window, out, hyper_prior, dpd_values):
    """Loom::dump (src/loom.cc:91-113) for one chain."""
    assigned_pos, n_assigned, order = window
    n_kinds, f2k = engine.get_kinds()
    specs, aux, kind_py, topology = engine.get_hypers()
    v = model_file.view
    model = Model(model_file.model().features, f2k, kind_py, topology)
    for i, f in enumerate(model.features):          # hypers as they stand on the device
        s = specs[i]
        if f["type"] == "bb":
            f.update(alpha=s.hp[0], beta=s.hp[1])
        elif f["type"] == "dd":
            f["alphas"] = aux[s.aux_off:s.aux_off + s.dim].tolist()
        elif f["type"] == "dpd":
            f.update(gamma=s.hp[0], alpha=s.hp[1], beta0=s.hp[2], betas=aux[s.aux_off:s.aux_off + s.dim].tolist())
        elif f["type"] == "gp":
            f.update(alpha=s.hp[0], inv_beta=s.hp[1])
        else:
            f.update(mu=s.hp[0], kappa=s.hp[1], sigmasq=s.hp[2], nu=s.hp[3])
    out_file = pbio.ModelFile.create(model, hyper_prior, dpd_values if v.n_dpd else None)
    if out.get("model"):
        out_file.write(out["model"])
    R = len(rowids)
    stream = (assigned_pos + np.arange(n_assigned)) % R              # FIFO order of the assigned window
    rows_of = stream if order is None else np.asarray(order)[stream]
    groupids = np.zeros((n_assigned, n_kinds), np.uint32)
    if out.get("groups"):
        os.makedirs(out["groups"], exist_ok=True)
    for k in range(n_kinds):
        counts, ints, flts = engine.get_groups(k)
        sorted_ids = pbio.sorted_groupids(counts)                                      # cross_cat.cc:212-236
        if out.get("groups"):
            pbio.groups_write(pbio.mixture_path(out["groups"], k), out_file, model.kind_features(k), counts, ints,
                              flts, sorted_ids)
        to_sorted = np.full(len(counts), _abi.UNASSIGNED, np.uint32)
        to_sorted[sorted_ids] = np.arange(len(sorted_ids), dtype=np.uint32)
        packed = engine.get_assignments(k)[rows_of]
        if n_assigned and (packed == _abi.UNASSIGNED).any():
            raise _abi.LoomError("bad id: an assigned row has no group in kind %d" % k)
        groupids[:, k] = to_sorted[packed]
    if out.get("assign"):
        pbio.assign_write(out["assign"], np.asarray(rowids)[rows_of], groupids)


def infer(rows_in, model_in, samples_out, config=None, tares_in=None, seeds=None, shuffle=True, abi=None, device=0):
    """Infer len(samples_out) samples of one dataset on one device.

    rows_in      rows.pbs[.gz] (the ingest's `diffs`); tares_in the matching tares file or None
    model_in     one init model path for every sample, or a list with one per sample
                 (loom.tasks.infer_one writes a fresh generate_init per seed)
    samples_out  per sample {"model": path, "groups": dir, "assign": path, "config": path} (any may be missing;
                 "config" receives the sample's Config with its own seed, what loom.tasks.infer_one leaves in the
                 store and loom_query reads back, loom/tasks.py:217-227)
    config       nested dict like loom/config.py DEFAULTS, or a config.pb[.gz] path
    seeds        per-sample seeds (default 0, 1, ...): the chain's rng and, with shuffle=True, its row order
    """
    abi = abi or _abi.load_cuda()                  # no CPU fallback: raises when the CUDA library is missing
    n = len(samples_out)
    seeds = list(range(n)) if seeds is None else list(seeds)
    if isinstance(config, (str, os.PathLike)):
        config = pbio.config_read(config).as_dict()
    model_paths = [model_in] * n if isinstance(model_in, (str, os.PathLike)) else list(model_in)
    files = [pbio.ModelFile.read(p) for p in model_paths]
    # DPD dictionaries are completed from the rows (loom's init models start with empty ones); every
    # sample then adopts the same value order, with betas broken off its own stick
    rows = pbio.rows_read(rows_in, files[0], open_dictionaries=True, seed=seeds[0])
    tare = (None, None)
    if tares_in:                                   # the tare row's own DPD values are in no row's pos
        tare_observed, tare_values, _ = pbio.tares_read(tares_in, files[0], open_dictionaries=True, seed=seeds[0])
        tare = (tare_observed, tare_values)
    for c in range(1, n):
        files[c].adopt_dictionaries(files[0], seeds[c])
    R = rows.n_rows
    empty = int(((config or {}).get("kernels", {}).get("cat", {}) or {}).get("empty_group_count", 1))
    engines = []
    for c in range(n):
        e = Engine(abi, device=device)
        e.set_model(files[c].model(empty))
        if c == 0:
            rows.upload(e, *tare)
        else:
            e.share_rows(engines[0])
        engines.append(e)
    orders = None
    if shuffle:                                    # loom.runner.shuffle(seed = sample seed), loom/tasks.py:228-233
        orders = np.stack([np.random.default_rng(s).permutation(R).astype(np.uint32) for s in seeds])
    hyper_prior = files[0].hyper_prior()
    extra_passes = ((config or {}).get("schedule", {}) or {}).get("extra_passes", 500.0)
    if extra_passes > 0:
        infer_multi_pass(engines, seeds, config, hyper_prior, orders=orders)
    else:
        infer_single_pass(engines, seeds, orders=orders)
    for c, e in enumerate(engines):
        if samples_out[c].get("config"):
            sample_config = pbio.config_from_dict(config)
            sample_config.seed = seeds[c]
            pbio.config_write(samples_out[c]["config"], sample_config)
        _dump_sample(e, files[c], rows.rowids, (0, R, None if orders is None else orders[c]), samples_out[c],
                     files[c].hyper_prior(), files[c].dpd_values())
    return engines


# ---------------------------------------------------------------------------------------------
# The ingest passes in process (loom.runner.tare / sparsify / shuffle, loom/runner.py:120-194): same arguments as the
# `loom_tare` / `loom_sparsify` / `loom_shuffle` processes, on the device engine through the C ABI.
def _schema_engine(schema_row_in, abi, device):
    mf = pbio.ModelFile.from_schema_row(schema_row_in)
    e = Engine(abi or _abi.load_cuda(), device=device)
    e.set_model(mf.model())
    return mf, e


def tare(schema_row_in, rows_in, tares_out, abi=None, device=0):
    """loom.runner.tare: the tare row of a dataset (the mode of every boolean / count column that more than half of
    the rows hold), written to `tares_out` -- or an empty file when no column qualifies"""
    mf, e = _schema_engine(schema_row_in, abi, device)
    rows = pbio.rows_read(rows_in, mf)
    if rows.row_has_tare.any():
        raise _abi.LoomError("row is already sparsified")
    rows.upload(e)
    tare_observed, tare_values = e.make_tare()
    pbio.tares_write(tares_out, mf, tare_observed, tare_values)
    return tare_observed, tare_values


def sparsify(schema_row_in, tares_in, rows_in, rows_out, abi=None, device=0):
    """loom.runner.sparsify: every row as Diff{pos, neg, tares: [0]} against the tare row of `tares_in`"""
    if str(rows_in) == str(rows_out):
        raise _abi.LoomError("in-place sparsify is not supported")
    mf, e = _schema_engine(schema_row_in, abi, device)
    tare_observed, tare_values, _ = pbio.tares_read(tares_in, mf)
    rows = pbio.rows_read(rows_in, mf)
    rows.upload(e)
    d = e.sparsify(tare_observed, tare_values)
    tared = np.full(rows.n_rows, 1 if tare_observed.any() else 0, np.uint8)
    out = pbio.Rows(rows.rowids, tared, d["pos_ptr"], d["pos_feature"], d["pos_value"], d["neg_ptr"], d["neg_feature"])
    pbio.rows_write(rows_out, mf, out, tare_values=tare_values)


def shuffle(rows_in, rows_out, seed=0, target_mem_bytes=4e9):
    """loom.runner.shuffle"""
    pbio.shuffle_stream(rows_in, rows_out, seed, target_mem_bytes)
