"""Host-side mirror of loom's engine objects over the C ABI.

`Engine` plays the part of `loom::Loom` (src/loom.hpp:43-135): it owns one
CrossCat + Assignments on one device and exposes the kernels by the reference's
names.  The same class drives the CUDA library (product) and, in tests only, the
CPU oracle -- the `abi` argument decides; nothing here falls back from one to
the other.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import (FeatureSpec, HyperPrior, KindInfo, LB_BB, LB_DD, LB_DPD, LB_GP, LB_NICH,
                   UNASSIGNED, ptr)

_u8p, _u32p, _f32p, _f64p = _abi._u8p, _abi._u32p, _abi._f32p, _abi._f64p


class Model:
    """CrossCat model: schema + hypers + kind structure (src/cross_cat.hpp:43-75).

    features: list of dicts {"type": "bb"|"dd"|"dpd"|"gp"|"nich", ...hypers} in
    loom's global feature order (bb < dd < dpd < gp < nich, dd by dim).
    """

    def __init__(self, features, featureid_to_kindid, kind_py=None, topology=(2.0, 0.1),
                 empty_group_count=1):
        self.features = [dict(f) for f in features]
        self.f2k = np.asarray(featureid_to_kindid, dtype=np.uint32).copy()
        assert len(self.f2k) == len(self.features)
        n_kinds = int(self.f2k.max()) + 1 if kind_py is None else len(kind_py)
        if kind_py is None:
            kind_py = [(2.0, 0.1)] * n_kinds  # loom/generate.py:41 CLUSTERING
        self.kind_py = np.asarray(kind_py, dtype=np.float32).reshape(-1, 2).copy()
        self.topology = np.asarray(topology, dtype=np.float32).copy()
        self.empty_group_count = int(empty_group_count)

    @property
    def n_features(self):
        return len(self.features)

    @property
    def n_kinds(self):
        return len(self.kind_py)

    def specs(self):
        F = len(self.features)
        specs = (FeatureSpec * F)()
        aux = []
        for i, f in enumerate(self.features):
            t = _abi.TYPE_NAMES[f["type"]]
            specs[i].type = t
            specs[i].dim = 0
            specs[i].aux_off = 0
            hp = [0.0] * 4
            if t == LB_BB:
                hp[:2] = [f["alpha"], f["beta"]]
            elif t == LB_DD:
                specs[i].dim = len(f["alphas"])
                specs[i].aux_off = len(aux)
                aux.extend(f["alphas"])
            elif t == LB_DPD:
                specs[i].dim = len(f["betas"])
                specs[i].aux_off = len(aux)
                aux.extend(f["betas"])
                hp[:3] = [f["gamma"], f["alpha"], f["beta0"]]
            elif t == LB_GP:
                hp[:2] = [f["alpha"], f["inv_beta"]]
            elif t == LB_NICH:
                hp[:4] = [f["mu"], f["kappa"], f["sigmasq"], f["nu"]]
            for j in range(4):
                specs[i].hp[j] = hp[j]
        return specs, np.asarray(aux if aux else [0.0], dtype=np.float32), len(aux)

    def n8(self):
        return sum(1 for f in self.features if f["type"] in ("bb", "dd"))

    def kind_features(self, k):
        return [i for i in range(self.n_features) if self.f2k[i] == k]

    @staticmethod
    def int_rows(f):
        t = f["type"]
        return {"bb": 2, "gp": 2, "nich": 1}.get(t) or (len(f["alphas"] if t == "dd" else f["betas"]) + 1)

    @staticmethod
    def flt_rows(f):
        return {"gp": 1, "nich": 2}.get(f["type"], 0)


class RowBlock:
    """Columnar row block (layout in include/loom_b200.h)."""

    def __init__(self, n_rows, cat8, wide, mask):
        self.n_rows = int(n_rows)
        self.cat8 = np.ascontiguousarray(cat8, dtype=np.uint8)
        self.wide = np.ascontiguousarray(wide, dtype=np.uint32)
        self.mask = np.ascontiguousarray(mask, dtype=np.uint32)

    @staticmethod
    def from_table(model, table):
        """table: list of rows, each a list of python values or None (the form
        loom/test/test_posterior_enum.py:584-618 generates)."""
        R, F, F8 = len(table), model.n_features, model.n8()
        cat8 = np.zeros((max(F8, 1), R), np.uint8)
        wide = np.zeros((max(F - F8, 1), R), np.uint32)
        obs = np.zeros((F, R), bool)
        for r, row in enumerate(table):
            for f, v in enumerate(row):
                if v is None:
                    continue
                obs[f, r] = True
                if f < F8:
                    cat8[f, r] = int(v)
                elif model.features[f]["type"] == "nich":
                    wide[f - F8, r] = np.float32(v).view(np.uint32)
                else:
                    wide[f - F8, r] = int(v)
        return RowBlock(R, cat8[:F8] if F8 else cat8[:0], wide[:F - F8] if F > F8 else wide[:0],
                        pack_mask(obs))

    def observed(self):
        F = self.mask.shape[0]
        bits = np.unpackbits(self.mask.view(np.uint8).reshape(F, -1), axis=1, bitorder="little")
        return bits[:, :self.n_rows].astype(bool)

    def cells_per_feature(self):
        """observed cells of every feature: popcount of the mask rows (padding bits are zero)"""
        return np.bitwise_count(self.mask).sum(axis=1, dtype=np.int64)

    def n_cells(self, features=None):
        per = self.cells_per_feature()
        if features is not None:
            per = per[list(features)]
        return int(per.sum())


def pack_mask(obs):
    """bool [F][R] -> uint32 [F][ceil(R/32)], bit r&31 of word r>>5."""
    F, R = obs.shape
    W = (R + 31) // 32
    padded = np.zeros((F, W * 32), np.uint8)
    padded[:, :R] = obs
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(F, W).copy()


def sweep_actions(n_rows, add=True, remove=False, start=0):
    """Action list for one pass in stream order.  add only: greedy ADD pass
    (Loom::infer_single_pass, src/loom.cc:112-137); remove+add: the
    posterior_enum / extra-pass refresh, REMOVE row then ADD row
    (src/loom.cc:502-505)."""
    rows = np.arange(start, start + n_rows, dtype=np.uint32)
    if add and remove:
        a = np.empty(2 * n_rows, np.uint32)
        a[0::2] = (rows << 1) | 1
        a[1::2] = rows << 1
        return a
    return ((rows << 1) | (0 if add else 1)).astype(np.uint32)


def cat_sweep_multi(engines, actions, seeds, action_offset=0):
    """The same action list on several independent chains (loom's samples, loom/tasks.py:173-194)
    in one call; chain c draws with seeds[c]."""
    if not engines:
        return
    actions = np.ascontiguousarray(actions, dtype=np.uint32)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    assert len(seeds) == len(engines)
    hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
    a = engines[0].abi
    a.check(a.cat_sweep_multi(hs, len(engines), ptr(actions, _u32p), len(actions),
                              seeds.ctypes.data_as(C.POINTER(C.c_uint64)), action_offset))


def cat_sweep_streams(engines, actions, seeds, action_offset=0):
    """One launch, a row order per chain: actions is [n_chains][n_actions] (loom shuffles the rows
    separately for every sample, loom/tasks.py:228-233; the chains still share one row block)."""
    if not engines:
        return
    actions = np.ascontiguousarray(actions, dtype=np.uint32)
    assert actions.ndim == 2 and actions.shape[0] == len(engines)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    assert len(seeds) == len(engines)
    hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
    a = engines[0].abi
    a.check(a.cat_sweep_streams(hs, len(engines), ptr(actions, _u32p), actions.shape[1],
                                seeds.ctypes.data_as(C.POINTER(C.c_uint64)), action_offset))


def query_score(engines, row_begin=0, row_end=None):
    """QueryServer::call(Score) over the latent samples `engines` (src/query_server.cc:189-231):
    log_sum_exp over samples of the per-sample score, minus log(#samples)."""
    per = np.stack([e.query_score(row_begin, row_end).astype(np.float64) for e in engines])
    m = per.max(axis=0)
    return m + np.log(np.exp(per - m).sum(axis=0)) - np.log(len(engines))


class Engine:
    def __init__(self, abi, device=0):
        self.abi = abi
        self.h = C.c_void_p()
        abi.check(abi.create(device, C.byref(self.h)))
        self.model = None
        self.n_rows = 0
        self._keep = []

    def close(self):
        if self.h:
            self.abi.destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- CrossCat::model_load + mixture_init_unobserved
    def set_model(self, model):
        specs, aux, n_aux = model.specs()
        a = self.abi
        a.check(a.set_model(self.h, model.n_features, specs, ptr(aux, _f32p), n_aux,
                            model.n_kinds, ptr(model.f2k, _u32p), ptr(model.kind_py, _f32p),
                            ptr(model.topology, _f32p), model.empty_group_count))
        self.model = model

    def set_rows(self, rows):
        a = self.abi
        a.check(a.set_rows(self.h, rows.n_rows, ptr(rows.cat8, _u8p), ptr(rows.wide, _u32p),
                           ptr(rows.mask, _u32p)))
        self.n_rows = rows.n_rows

    def reload_rows(self, rows):
        """stream the same block in again from host buffers (the reference re-reads its rows file every pass);
        assignments and statistics are kept"""
        a = self.abi
        a.check(a.reload_rows(self.h, rows.n_rows, ptr(rows.cat8, _u8p), ptr(rows.wide, _u32p), ptr(rows.mask, _u32p)))

    def generate_rows(self, n_rows, density, seed):
        """Loom::generate's generate_rows: synthesise rows from the model; the engine then holds them, assigned"""
        self.abi.check(self.abi.generate_rows(self.h, n_rows, density, seed))
        self.n_rows = int(n_rows)

    def get_rows(self):
        """the loaded block as a RowBlock"""
        F, F8, R = self.model.n_features, self.model.n8(), self.n_rows
        cat8 = np.zeros((max(F8, 1), R), np.uint8)
        wide = np.zeros((max(F - F8, 1), R), np.uint32)
        mask = np.zeros((F, (R + 31) // 32), np.uint32)
        self.abi.check(self.abi.get_rows(self.h, ptr(cat8, _u8p), ptr(wide, _u32p), ptr(mask, _u32p)))
        return RowBlock(R, cat8[:F8] if F8 else cat8[:0], wide[:F - F8] if F > F8 else wide[:0], mask)

    def share_rows(self, owner):
        """read the owner's row block (loom's samples share one rows file)"""
        self.abi.check(self.abi.share_rows(self.h, owner.h))
        self.n_rows = owner.n_rows
        self._rows_owner = owner       # keep it alive

    def set_rows_diff(self, n_rows, tare_observed, tare_values, row_has_tare, pos_ptr, pos_feature, pos_value,
                      neg_ptr, neg_feature):
        """Tare-compressed rows (ProductValue::Diff), see loom_b200/tare.py."""
        c = lambda x, dt: None if x is None else np.ascontiguousarray(x, dtype=dt)
        to, tv, rh = c(tare_observed, np.uint8), c(tare_values, np.uint32), c(row_has_tare, np.uint8)
        pp, pf, pv = c(pos_ptr, np.uint64), c(pos_feature, np.uint32), c(pos_value, np.uint32)
        npt, nf = c(neg_ptr, np.uint64), c(neg_feature, np.uint32)
        a = self.abi
        a.check(a.set_rows_diff(self.h, n_rows, ptr(to, _u8p), ptr(tv, _u32p), ptr(rh, _u8p), ptr(pp, _abi._u64p),
                                ptr(pf, _u32p), ptr(pv, _u32p), ptr(npt, _abi._u64p), ptr(nf, _u32p)))
        self.n_rows = int(n_rows)

    def n_kinds(self):
        n = C.c_int32()
        self.abi.check(self.abi.get_kinds(self.h, C.byref(n), None))
        return n.value

    def get_kinds(self):
        n = C.c_int32()
        f2k = np.zeros(self.model.n_features, np.uint32)
        self.abi.check(self.abi.get_kinds(self.h, C.byref(n), ptr(f2k, _u32p)))
        return n.value, f2k

    # -- CatKernel::add_row / remove_row over an action list
    def cat_sweep(self, actions, seed, action_offset=0, kind_mask=None, debug=False,
                  debug_stride=0):
        actions = np.ascontiguousarray(actions, dtype=np.uint32)
        K = self.n_kinds()
        picks = scores = None
        if debug:
            picks = np.full((len(actions), K), UNASSIGNED, np.uint32)
            if debug_stride:
                scores = np.full((len(actions), K, debug_stride), np.nan, np.float32)
        km = None if kind_mask is None else np.ascontiguousarray(kind_mask, dtype=np.uint8)
        a = self.abi
        a.check(a.cat_sweep(self.h, ptr(actions, _u32p), len(actions), seed, action_offset,
                            ptr(km, _u8p), ptr(picks, _u32p), ptr(scores, _f32p), debug_stride))
        return picks, scores

    def get_assignments(self, kind):
        out = np.empty(self.n_rows, np.uint32)
        self.abi.check(self.abi.get_assignments(self.h, kind, ptr(out, _u32p)))
        return out

    def set_assignments(self, kind, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        assert len(packed) == self.n_rows
        self.abi.check(self.abi.set_assignments(self.h, kind, ptr(packed, _u32p)))

    def kind_info(self, kind):
        info = KindInfo()
        self.abi.check(self.abi.kind_info(self.h, kind, C.byref(info)))
        return info

    def get_groups(self, kind):
        info = self.kind_info(kind)
        G = info.n_groups
        counts = np.zeros(G, np.uint32)
        ints = np.zeros((max(info.n_int_rows, 1), G), np.uint32)
        flts = np.zeros((max(info.n_flt_rows, 1), G), np.float64)
        self.abi.check(self.abi.get_groups(self.h, kind, ptr(counts, _u32p), ptr(ints, _u32p),
                                           ptr(flts, _f64p)))
        return counts, ints[:info.n_int_rows], flts[:info.n_flt_rows]

    def set_groups(self, kind, n_groups, counts=None, ints=None, flts=None):
        def prep(x, dt):
            return None if x is None else np.ascontiguousarray(x, dtype=dt)
        counts, ints, flts = prep(counts, np.uint32), prep(ints, np.uint32), prep(flts, np.float64)
        self.abi.check(self.abi.set_groups(self.h, kind, n_groups, ptr(counts, _u32p),
                                           ptr(ints, _u32p), ptr(flts, _f64p)))

    # -- ProductMixture::score_value, batched
    def score_rows(self, kind, row_begin=0, row_end=None, fetch=True):
        row_end = self.n_rows if row_end is None else row_end
        G = self.kind_info(kind).n_groups
        out = np.empty((row_end - row_begin, G), np.float32) if fetch else None
        self.abi.check(self.abi.score_rows(self.h, kind, row_begin, row_end, ptr(out, _f32p)))
        return out

    def score_rows_diff(self, kind, row_begin=0, row_end=None, fetch=True):
        row_end = self.n_rows if row_end is None else row_end
        G = self.kind_info(kind).n_groups
        out = np.empty((row_end - row_begin, G), np.float32) if fetch else None
        self.abi.check(self.abi.score_rows_diff(self.h, kind, row_begin, row_end, ptr(out, _f32p)))
        return out

    def query_score(self, row_begin=0, row_end=None):
        """log predictive probability of every loaded row under this sample (QueryServer Score)"""
        row_end = self.n_rows if row_end is None else row_end
        out = np.empty(row_end - row_begin, np.float32)
        self.abi.check(self.abi.query_score(self.h, row_begin, row_end, ptr(out, _f32p)))
        return out

    # -- loom::Differ (src/differ.cc): the tare / sparsify passes over the loaded block
    def make_tare(self, dpd_offsets=None, dpd_values=None, reset=True):
        """Differ::add_rows + _make_tare on the loaded block -> (tare_observed uint8 [F], tare_values uint32 [F]);
        reset=False keeps adding to the summaries of earlier blocks (windows of one stream)"""
        i32p = C.POINTER(C.c_int32)
        off = None if dpd_offsets is None else np.ascontiguousarray(dpd_offsets, np.int32)
        val = None if dpd_values is None else np.ascontiguousarray(dpd_values, np.uint32)
        if reset:
            self.abi.check(self.abi.differ_reset(self.h))
        self.abi.check(self.abi.differ_add_rows(self.h, ptr(off, i32p), ptr(val, _u32p)))
        F = self.model.n_features
        obs, values = np.zeros(F, np.uint8), np.zeros(F, np.uint32)
        self.abi.check(self.abi.differ_make_tare(self.h, ptr(off, i32p), ptr(val, _u32p), ptr(obs, _abi._u8p), ptr(values, _u32p)))
        return obs, values

    def sparsify(self, tare_observed, tare_values):
        """Differ::compress_rows on the loaded block -> dict of the arrays set_rows_diff takes"""
        obs = np.ascontiguousarray(tare_observed, np.uint8)
        val = np.ascontiguousarray(tare_values, np.uint32)
        n_pos, n_neg = C.c_uint64(), C.c_uint64()
        self.abi.check(self.abi.differ_compress_rows(self.h, ptr(obs, _abi._u8p), ptr(val, _u32p), C.byref(n_pos), C.byref(n_neg)))
        R = self.n_rows
        pos_ptr, neg_ptr = np.zeros(R + 1, np.uint64), np.zeros(R + 1, np.uint64)
        pos_f, pos_v = np.zeros(n_pos.value, np.uint32), np.zeros(n_pos.value, np.uint32)
        neg_f = np.zeros(n_neg.value, np.uint32)
        self.abi.check(self.abi.differ_fetch(self.h, ptr(pos_ptr, _abi._u64p), ptr(pos_f, _u32p), ptr(pos_v, _u32p),
                                             ptr(neg_ptr, _abi._u64p), ptr(neg_f, _u32p)))
        return dict(n_rows=R, tare_observed=obs, tare_values=val, row_has_tare=None, pos_ptr=pos_ptr, pos_feature=pos_f,
                    pos_value=pos_v, neg_ptr=neg_ptr, neg_feature=neg_f)

    def query_sample(self, row, to_sample, n_samples, seed, draw_offset=0, groups=False):
        """n_samples joint draws of the features `to_sample` given the loaded row `row` under this sample
        (the per-latent half of QueryServer Sample): uint32 [n_samples][len(to_sample)] raw cell values
        (NICH: float bits; DPD_OTHER for a value outside a DPD dictionary), and with groups=True also the
        packed group drawn per kind [n_samples][n_kinds]."""
        to_sample = np.ascontiguousarray(to_sample, np.uint32)
        values = np.empty((n_samples, len(to_sample)), np.uint32)
        picked = np.empty((n_samples, self.n_kinds()), np.uint32) if groups else None
        self.abi.check(self.abi.query_sample(self.h, row, ptr(to_sample, _u32p), len(to_sample), n_samples, seed,
                                             draw_offset, ptr(values, _u32p), ptr(picked, _u32p)))
        return (values, picked) if groups else values

    def accumulate_rows(self, kind=-1, row_begin=0, row_end=None, sign=1):
        row_end = self.n_rows if row_end is None else row_end
        self.abi.check(self.abi.accumulate_rows(self.h, kind, row_begin, row_end, sign))

    def score_data(self):
        out = C.c_double()
        self.abi.check(self.abi.score_data(self.h, C.byref(out)))
        return out.value

    # -- KindKernel
    def kind_proposer_score(self, fetch=True):
        K = self.n_kinds()
        out = np.empty((self.model.n_features, K), np.float32) if fetch else None
        self.abi.check(self.abi.kind_proposer_score(self.h, ptr(out, _f32p)))
        return out

    def proposer_accumulate(self, row_begin=0, row_end=None):
        row_end = self.n_rows if row_end is None else row_end
        self.abi.check(self.abi.proposer_accumulate(self.h, row_begin, row_end))

    def proposer_tables(self, kind):
        """(ints_ptr, n_ints, flts_ptr, n_flts): flat buffers to all-reduce (device pointers for CUDA)."""
        pi, pf = C.c_void_p(), C.c_void_p()
        ni, nf = C.c_uint64(), C.c_uint64()
        self.abi.check(self.abi.proposer_tables(self.h, kind, C.byref(pi), C.byref(ni), C.byref(pf), C.byref(nf)))
        return pi.value or 0, ni.value, pf.value or 0, nf.value

    def proposer_finish(self, fetch=True):
        K = self.n_kinds()
        out = np.empty((self.model.n_features, K), np.float32) if fetch else None
        self.abi.check(self.abi.proposer_finish(self.h, ptr(out, _f32p)))
        return out

    def kind_run(self, iterations, seed):
        f2k = np.zeros(self.model.n_features, np.uint32)
        n = C.c_int32()
        self.abi.check(self.abi.kind_run(self.h, iterations, seed, ptr(f2k, _u32p), C.byref(n)))
        return f2k, n.value

    def prepare_featureless_kinds(self, count, seed):
        """start the next init_featureless_kinds(count, seed)'s host-side seating in the background (under a sweep)"""
        self.abi.check(self.abi.prepare_featureless_kinds(self.h, count, seed))

    def init_featureless_kinds(self, count, seed):
        self.abi.check(self.abi.init_featureless_kinds(self.h, count, seed))

    # -- HyperKernel
    def set_hyper_prior(self, grids):
        """grids: dict like loom/hyperprior.py DEFAULTS, e.g.
        {"topology": [(a, d), ...], "bb": {"alpha": [...], "beta": [...]}, ...}"""
        hp = HyperPrior()
        keep = []

        def put(name, values, width=1):
            arr = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
            keep.append(arr)
            setattr(hp, name, ptr(arr, _f32p) if arr.size else None)
            setattr(hp, "n_" + name, arr.size // width)

        for name in ("topology", "clustering"):
            pts = grids.get(name, [])
            put(name, [[p["alpha"], p["d"]] if isinstance(p, dict) else list(p) for p in pts], 2)
        for t, params in (("bb", ("alpha", "beta")), ("dd", ("alpha",)), ("dpd", ("gamma", "alpha")),
                          ("gp", ("alpha", "inv_beta")), ("nich", ("mu", "kappa", "sigmasq", "nu"))):
            for p in params:
                put("%s_%s" % (t, p), grids.get(t, {}).get(p, []))
        self.abi.check(self.abi.set_hyper_prior(self.h, C.byref(hp)))
        self._keep = keep

    def hyper_run(self, seed, task_mask=None):
        """HyperKernel::run; task_mask [1 + K + F] (topology | kind clusterings | features) runs a share of the tasks"""
        if task_mask is None:
            self.abi.check(self.abi.hyper_run(self.h, seed))
        else:
            m = np.ascontiguousarray(task_mask, np.uint8)
            assert len(m) == 1 + self.n_kinds() + self.model.n_features
            self.abi.check(self.abi.hyper_run_tasks(self.h, seed, ptr(m, _u8p)))

    def set_hypers(self, specs, aux, kind_py, topology):
        """install hypers chosen elsewhere (the arrays get_hypers returns)"""
        full_aux = np.zeros(max(len(aux), 1), np.float32)
        full_aux[:len(aux)] = aux
        kind_py = np.ascontiguousarray(kind_py, np.float32)
        topology = np.ascontiguousarray(topology, np.float32)
        self.abi.check(self.abi.set_hypers(self.h, specs, ptr(full_aux, _f32p), ptr(kind_py, _f32p), ptr(topology, _f32p)))

    def get_hypers(self):
        F = self.model.n_features
        specs = (FeatureSpec * F)()
        _, aux, n_aux = self.model.specs()
        K = self.n_kinds()
        kind_py = np.zeros((K, 2), np.float32)
        topo = np.zeros(2, np.float32)
        self.abi.check(self.abi.get_hypers(self.h, specs, ptr(aux, _f32p), ptr(kind_py, _f32p),
                                           ptr(topo, _f32p)))
        return specs, aux[:n_aux], kind_py, topo

    # -- timing
    def sync(self):
        self.abi.check(self.abi.sync(self.h))

    def timer_start(self):
        self.abi.check(self.abi.timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.abi.check(self.abi.timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False, n=8):
        out = np.zeros(n, np.uint64)
        self.abi.check(self.abi.launch_count(self.h, ptr(out, _abi._u64p), n, int(reset)))
        return out

    def kernel_ms(self, reset=False, n=8):
        out = np.zeros(n, np.float64)
        self.abi.check(self.abi.kernel_ms(self.h, ptr(out, _f64p), n, int(reset)))
        return out
