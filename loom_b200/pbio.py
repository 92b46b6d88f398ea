"""ctypes binding of include/loom_b200_io.h: loom's protobuf files (config, model, groups,
assignments, rows, tares, checkpoint, log) <-> the arrays the engine's C ABI takes.

The decoding and encoding are native (loom_b200/io/lb_io.cc); this module only moves
pointers.  It is the file half of the `loom.tasks.infer` drop-in.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import FeatureSpec, HyperPrior, LoomError
from .engine import Model

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IO_LIB = os.path.join(_REPO, "loom_b200", "io", "libloomb200_io.so")

_u8p, _u32p, _u64p, _f32p, _f64p = _abi._u8p, _abi._u32p, _abi._u64p, _abi._f32p, _abi._f64p
_i32p = C.POINTER(C.c_int32)


class Config(C.Structure):
    """protobuf::Config (src/schema.proto:167-229)"""
    _fields_ = [("seed", C.c_uint64),
                ("extra_passes", C.c_float), ("small_data_size", C.c_float), ("big_data_size", C.c_float),
                ("max_reject_iters", C.c_uint32), ("checkpoint_period_sec", C.c_float),
                ("cat_empty_group_count", C.c_uint32), ("cat_row_queue_capacity", C.c_uint32),
                ("cat_parser_threads", C.c_uint32),
                ("hyper_run", C.c_int32), ("hyper_parallel", C.c_int32),
                ("kind_iterations", C.c_uint32), ("kind_empty_kind_count", C.c_uint32),
                ("kind_row_queue_capacity", C.c_uint32), ("kind_parser_threads", C.c_uint32),
                ("kind_score_parallel", C.c_int32),
                ("posterior_enum_sample_count", C.c_uint32), ("posterior_enum_sample_skip", C.c_uint32),
                ("generate_row_count", C.c_uint64), ("generate_density", C.c_float),
                ("generate_sample_skip", C.c_uint32), ("target_mem_bytes", C.c_float),
                ("has_query", C.c_int32), ("query_parallel", C.c_int32)]

    def as_dict(self):
        """the nested dict shape of loom/config.py DEFAULTS (what loom_b200.infer takes)"""
        return {"seed": self.seed, "target_mem_bytes": self.target_mem_bytes,
                "schedule": {"extra_passes": self.extra_passes, "small_data_size": self.small_data_size,
                             "big_data_size": self.big_data_size, "max_reject_iters": self.max_reject_iters,
                             "checkpoint_period_sec": self.checkpoint_period_sec},
                "kernels": {"cat": {"empty_group_count": self.cat_empty_group_count,
                                    "row_queue_capacity": self.cat_row_queue_capacity,
                                    "parser_threads": self.cat_parser_threads},
                            "hyper": {"run": bool(self.hyper_run), "parallel": bool(self.hyper_parallel)},
                            "kind": {"iterations": self.kind_iterations, "empty_kind_count": self.kind_empty_kind_count,
                                     "row_queue_capacity": self.kind_row_queue_capacity,
                                     "parser_threads": self.kind_parser_threads,
                                     "score_parallel": bool(self.kind_score_parallel)}},
                "posterior_enum": {"sample_count": self.posterior_enum_sample_count,
                                   "sample_skip": self.posterior_enum_sample_skip},
                "generate": {"row_count": self.generate_row_count, "density": self.generate_density,
                             "sample_skip": self.generate_sample_skip},
                "query": {"parallel": bool(self.query_parallel)}}


def config_from_dict(d=None):
    """the inverse of Config.as_dict over loom's defaults (loom/config.py:76-117 fill_in_defaults + protobuf_dump)"""
    c = config_defaults()
    d = d or {}
    sch, ker = d.get("schedule") or {}, d.get("kernels") or {}
    cat, hyp, kind = ker.get("cat") or {}, ker.get("hyper") or {}, ker.get("kind") or {}
    pe, gen, qry = d.get("posterior_enum") or {}, d.get("generate") or {}, d.get("query") or {}
    for field, src, key in (("seed", d, "seed"), ("target_mem_bytes", d, "target_mem_bytes"),
                            ("extra_passes", sch, "extra_passes"), ("small_data_size", sch, "small_data_size"),
                            ("big_data_size", sch, "big_data_size"), ("max_reject_iters", sch, "max_reject_iters"),
                            ("checkpoint_period_sec", sch, "checkpoint_period_sec"),
                            ("cat_empty_group_count", cat, "empty_group_count"),
                            ("cat_row_queue_capacity", cat, "row_queue_capacity"), ("cat_parser_threads", cat, "parser_threads"),
                            ("hyper_run", hyp, "run"), ("hyper_parallel", hyp, "parallel"),
                            ("kind_iterations", kind, "iterations"), ("kind_empty_kind_count", kind, "empty_kind_count"),
                            ("kind_row_queue_capacity", kind, "row_queue_capacity"),
                            ("kind_parser_threads", kind, "parser_threads"), ("kind_score_parallel", kind, "score_parallel"),
                            ("posterior_enum_sample_count", pe, "sample_count"), ("posterior_enum_sample_skip", pe, "sample_skip"),
                            ("generate_row_count", gen, "row_count"), ("generate_density", gen, "density"),
                            ("generate_sample_skip", gen, "sample_skip"), ("query_parallel", qry, "parallel")):
        if key in src and src[key] is not None:
            setattr(c, field, type(getattr(c, field))(src[key]))
    return c


class Checkpoint(C.Structure):
    """protobuf::Checkpoint (src/schema.proto:234-253)"""
    _fields_ = [("finished", C.c_int32), ("seed", C.c_uint64), ("tardis_iter", C.c_uint64),
                ("annealing_state", C.c_double), ("schedule_row_count", C.c_uint64), ("reject_iters", C.c_uint64),
                ("row_count", C.c_uint64), ("unassigned_pos", C.c_uint64), ("assigned_pos", C.c_uint64)]


class ModelView(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("specs", C.POINTER(FeatureSpec)), ("aux", _f32p), ("n_aux", C.c_int32),
                ("n_kinds", C.c_int32), ("featureid_to_kindid", _u32p), ("kind_py", _f32p),
                ("topology_py", C.c_float * 2), ("has_hyper_prior", C.c_int32), ("hyper_prior", HyperPrior),
                ("n_dpd", C.c_int32), ("dpd_offsets", _i32p), ("dpd_values", _u32p)]


class RowBlockView(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("rowids", _u64p), ("row_has_tare", _u8p),
                ("pos_ptr", _u64p), ("pos_feature", _u32p), ("pos_value", _u32p),
                ("neg_ptr", _u64p), ("neg_feature", _u32p), ("stream_rows", C.c_uint64), ("owner", C.c_void_p)]


class AssignBlock(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_kinds", C.c_int32), ("rowids", _u64p), ("groupids", _u32p),
                ("owner", C.c_void_p)]


class GroupBlock(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_int_rows", C.c_int32), ("n_flt_rows", C.c_int32),
                ("counts", _u32p), ("ints", _u32p), ("flts", _f64p), ("owner", C.c_void_p)]


class LogEntry(C.Structure):
    _fields_ = [("iter", C.c_uint32), ("topology_py", C.c_float * 2), ("n_kinds", C.c_int32), ("kind_py", _f32p),
                ("feature_counts", _u32p), ("category_counts", _u32p), ("score", C.c_float),
                ("kl_divergence", C.c_float), ("assigned_object_count", C.c_uint64),
                ("cat_total_time_usec", C.c_uint64), ("hyper_total_time_usec", C.c_uint64),
                ("kind_total_time_usec", C.c_uint64), ("kind_total_count", C.c_uint64),
                ("kind_change_count", C.c_uint64)]


_P = C.POINTER
_cs = C.c_char_p
_model = C.c_void_p
SIGNATURES = {   # must list every function of include/loom_b200_io.h
    "last_error": (C.c_char_p, []),
    "config_defaults": (None, [_P(Config)]),
    "config_read": (C.c_int, [_cs, _P(Config)]),
    "config_write": (C.c_int, [_cs, _P(Config)]),
    "checkpoint_read": (C.c_int, [_cs, _P(Checkpoint)]),
    "checkpoint_write": (C.c_int, [_cs, _P(Checkpoint)]),
    "model_read": (C.c_int, [_cs, _P(_model)]),
    "model_create": (C.c_int, [C.c_int32, _P(FeatureSpec), _f32p, C.c_int32, C.c_int32, _u32p, _f32p, _f32p,
                               _P(HyperPrior), _i32p, _u32p, _P(_model)]),
    "model_get": (C.c_int, [_model, _P(ModelView)]),
    "model_write": (C.c_int, [_model, _cs]),
    "model_free": (None, [_model]),
    "rows_read": (C.c_int, [_cs, _model, C.c_uint64, C.c_uint64, C.c_int32, _P(RowBlockView)]),
    "rows_free": (None, [_P(RowBlockView)]),
    "rows_read_open": (C.c_int, [_cs, _model, C.c_uint64, C.c_int32, _P(RowBlockView), _i32p]),
    "model_adopt_dictionaries": (C.c_int, [_model, _model, C.c_uint64]),
    "stream_stats": (C.c_int, [_cs, _u64p, _u32p]),
    "rows_write": (C.c_int, [_cs, _model, _P(RowBlockView)]),
    "rows_to_columns": (C.c_int, [_model, _P(RowBlockView), _u8p, _u32p, _u8p, _u32p, _u32p]),
    "tares_read": (C.c_int, [_cs, _model, _u8p, _u32p, _i32p]),
    "tares_write": (C.c_int, [_cs, _model, _u8p, _u32p]),
    "assign_read": (C.c_int, [_cs, _P(AssignBlock)]),
    "assign_free": (None, [_P(AssignBlock)]),
    "assign_write": (C.c_int, [_cs, C.c_uint64, C.c_int32, _u64p, _u32p]),
    "groups_read": (C.c_int, [_cs, _model, _u32p, C.c_int32, _P(GroupBlock)]),
    "groups_free": (None, [_P(GroupBlock)]),
    "groups_write": (C.c_int, [_cs, _model, _u32p, C.c_int32, C.c_int32, _u32p, _u32p, _f64p, _u32p, C.c_int32]),
    "log_open": (C.c_int, [_cs, C.c_int32, _P(C.c_void_p)]),
    "log_write": (C.c_int, [C.c_void_p, _P(LogEntry)]),
    "log_close": (None, [C.c_void_p]),
    "tares_read_open": (C.c_int, [_cs, _model, C.c_uint64, _u8p, _u32p, _i32p, _i32p]),
    "shuffle_stream": (C.c_int, [_cs, _cs, C.c_uint64, C.c_double]),
    "sorted_groupids": (C.c_int, [_u32p, C.c_int32, _u32p, _i32p]),
    "rows_writer_open": (C.c_int, [_cs, _P(C.c_void_p)]),
    "rows_writer_write": (C.c_int, [C.c_void_p, _model, C.c_void_p, _u32p]),
    "rows_writer_close": (C.c_int, [C.c_void_p]),
    "model_from_schema_row": (C.c_int, [_cs, _P(C.c_void_p)]),
    "query_open": (C.c_int, [_cs, _cs, _P(C.c_void_p)]),
    "query_next": (C.c_int, [C.c_void_p, _model, C.c_void_p, _i32p]),
    "query_respond": (C.c_int, [C.c_void_p, _model, C.c_void_p]),
    "query_close": (None, [C.c_void_p]),
    "samples_open": (C.c_int, [_cs, _P(C.c_void_p)]),
    "samples_write": (C.c_int, [C.c_void_p, C.c_int32, _i32p, _u32p, _i32p, _u64p, _u32p, C.c_float]),
    "samples_close": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib():
    """Bind libloomb200_io.so.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(IO_LIB):
            raise LoomError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % IO_LIB)
        handle = C.CDLL(IO_LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, "lbio_" + name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def _check(status):
    if status != 0:
        raise LoomError((lib().lbio_last_error() or b"?").decode(errors="replace"))


def _path(p):
    return os.fsencode(p)


def _arr(p, n, dtype):
    """copy n elements out of a library-owned array"""
    if not n:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(p, shape=(int(n),)).astype(dtype, copy=True)


# ------------------------------------------------------------------------------------------ config / checkpoint
def config_defaults():
    c = Config()
    lib().lbio_config_defaults(C.byref(c))
    return c


def config_read(path):
    c = Config()
    _check(lib().lbio_config_read(_path(path), C.byref(c)))
    return c


def config_write(path, config):
    _check(lib().lbio_config_write(_path(path), C.byref(config)))


def checkpoint_read(path):
    c = Checkpoint()
    _check(lib().lbio_checkpoint_read(_path(path), C.byref(c)))
    return c


def checkpoint_write(path, checkpoint):
    _check(lib().lbio_checkpoint_write(_path(path), C.byref(checkpoint)))


# ------------------------------------------------------------------------------------------ model
_GRID_KEYS = {"bb": ("alpha", "beta"), "dd": ("alpha",), "dpd": ("gamma", "alpha"), "gp": ("alpha", "inv_beta"),
              "nich": ("mu", "kappa", "sigmasq", "nu")}
_TYPE_OF = {v: k for k, v in _abi.TYPE_NAMES.items()}


class ModelFile:
    """A CrossCat model file in memory (the native lbio_model) and its `Model` view."""

    def __init__(self, handle):
        self.h = handle
        self.refresh()

    def refresh(self):
        """re-fetch the view (after the dictionaries grew)"""
        v = ModelView()
        _check(lib().lbio_model_get(self.h, C.byref(v)))
        self.view = v

    def adopt_dictionaries(self, other, seed):
        """take `other`'s DPD dictionaries; missing values get this model's own stick-breaking betas"""
        _check(lib().lbio_model_adopt_dictionaries(self.h, other.h, seed))
        self.refresh()

    def __del__(self):
        try:
            if self.h:
                lib().lbio_model_free(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def read(path):
        h = _model()
        _check(lib().lbio_model_read(_path(path), C.byref(h)))
        return ModelFile(h)

    @staticmethod
    def from_schema_row(path):
        """the stand-in model loom_tare / loom_sparsify decode rows with (only column counts are known)"""
        h = _model()
        _check(lib().lbio_model_from_schema_row(_path(path), C.byref(h)))
        return ModelFile(h)

    @staticmethod
    def create(model, hyper_prior=None, dpd_values=None):
        """from an engine `Model` (+ grids shaped like loom/hyperprior.py, + per-DPD-feature dictionaries)"""
        specs, aux, n_aux = model.specs()
        hp, keep = (None, None) if hyper_prior is None else grids_to_struct(hyper_prior)
        offs = vals = None
        if dpd_values is not None:
            offs = np.cumsum([0] + [len(d) for d in dpd_values]).astype(np.int32)
            vals = np.concatenate([np.asarray(d, np.uint32) for d in dpd_values] + [np.zeros(0, np.uint32)])
        h = _model()
        _check(lib().lbio_model_create(model.n_features, specs, _abi.ptr(aux, _f32p), n_aux, model.n_kinds,
                                       _abi.ptr(model.f2k, _u32p), _abi.ptr(model.kind_py, _f32p),
                                       _abi.ptr(model.topology, _f32p), None if hp is None else C.byref(hp),
                                       _abi.ptr(offs, _i32p), _abi.ptr(vals, _u32p), C.byref(h)))
        return ModelFile(h)

    def write(self, path):
        _check(lib().lbio_model_write(self.h, _path(path)))

    def model(self, empty_group_count=1):
        """the engine-side `Model` (what Engine.set_model takes)"""
        v = self.view
        aux = _arr(v.aux, v.n_aux, np.float32)
        feats = []
        for i in range(v.n_features):
            s = v.specs[i]
            t = _TYPE_OF[s.type]
            if t == "bb":
                feats.append({"type": t, "alpha": s.hp[0], "beta": s.hp[1]})
            elif t == "dd":
                feats.append({"type": t, "alphas": aux[s.aux_off:s.aux_off + s.dim].tolist()})
            elif t == "dpd":
                feats.append({"type": t, "gamma": s.hp[0], "alpha": s.hp[1], "beta0": s.hp[2],
                              "betas": aux[s.aux_off:s.aux_off + s.dim].tolist()})
            elif t == "gp":
                feats.append({"type": t, "alpha": s.hp[0], "inv_beta": s.hp[1]})
            else:
                feats.append({"type": t, "mu": s.hp[0], "kappa": s.hp[1], "sigmasq": s.hp[2], "nu": s.hp[3]})
        f2k = _arr(v.featureid_to_kindid, v.n_features, np.uint32)
        kind_py = _arr(v.kind_py, 2 * v.n_kinds, np.float32).reshape(-1, 2)
        return Model(feats, f2k, kind_py, (v.topology_py[0], v.topology_py[1]), empty_group_count)

    def hyper_prior(self):
        """grids as the dict Engine.set_hyper_prior takes, or None when the file has none"""
        v = self.view
        if not v.has_hyper_prior:
            return None
        hp = v.hyper_prior
        out = {"topology": _arr(hp.topology, 2 * hp.n_topology, np.float32).reshape(-1, 2).tolist(),
               "clustering": _arr(hp.clustering, 2 * hp.n_clustering, np.float32).reshape(-1, 2).tolist()}
        for t, params in _GRID_KEYS.items():
            out[t] = {p: _arr(getattr(hp, "%s_%s" % (t, p)), getattr(hp, "n_%s_%s" % (t, p)), np.float32).tolist()
                      for p in params}
        return out

    def dpd_values(self):
        v = self.view
        offs = _arr(v.dpd_offsets, v.n_dpd + 1, np.int32)
        vals = _arr(v.dpd_values, offs[-1] if len(offs) else 0, np.uint32)
        return [vals[offs[i]:offs[i + 1]] for i in range(v.n_dpd)]


def grids_to_struct(grids):
    hp = HyperPrior()
    keep = []

    def put(name, values, width=1):
        arr = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        keep.append(arr)
        setattr(hp, name, _abi.ptr(arr, _f32p) if arr.size else None)
        setattr(hp, "n_" + name, arr.size // width)

    for name in ("topology", "clustering"):
        pts = grids.get(name, [])
        put(name, [[p["alpha"], p["d"]] if isinstance(p, dict) else list(p) for p in pts], 2)
    for t, params in _GRID_KEYS.items():
        for p in params:
            put("%s_%s" % (t, p), grids.get(t, {}).get(p, []))
    hp._keep = keep
    return hp, keep


# ------------------------------------------------------------------------------------------ rows / tares
class Rows:
    """Decoded rows in the CSR form of lb_set_rows_diff (numpy copies)."""

    def __init__(self, rowids, row_has_tare, pos_ptr, pos_feature, pos_value, neg_ptr, neg_feature, stream_rows=0):
        self.rowids = np.ascontiguousarray(rowids, np.uint64)
        self.row_has_tare = np.ascontiguousarray(row_has_tare, np.uint8)
        self.pos_ptr = np.ascontiguousarray(pos_ptr, np.uint64)
        self.pos_feature = np.ascontiguousarray(pos_feature, np.uint32)
        self.pos_value = np.ascontiguousarray(pos_value, np.uint32)
        self.neg_ptr = np.ascontiguousarray(neg_ptr, np.uint64)
        self.neg_feature = np.ascontiguousarray(neg_feature, np.uint32)
        self.stream_rows = int(stream_rows)

    @property
    def n_rows(self):
        return len(self.rowids)

    @staticmethod
    def from_block(block, rowids=None):
        """a dense columnar `RowBlock` as untared rows: diff{pos = the row, neg = nothing}"""
        obs = block.observed()                                   # [F][R]
        F, R = obs.shape
        F8 = block.cat8.shape[0]
        vals = np.zeros((F, R), np.uint32)
        vals[:F8] = block.cat8
        vals[F8:] = block.wide
        ptr = np.zeros(R + 1, np.uint64)
        ptr[1:] = np.cumsum(obs.sum(axis=0))
        r_idx, f_idx = np.nonzero(obs.T)
        return Rows(np.arange(R) if rowids is None else rowids, np.zeros(R, np.uint8), ptr, f_idx,
                    vals[f_idx, r_idx], np.zeros(R + 1, np.uint64), np.zeros(0, np.uint32))

    def _view(self):
        v = RowBlockView()
        v.n_rows = self.n_rows
        v.rowids = _abi.ptr(self.rowids, _u64p)
        v.row_has_tare = _abi.ptr(self.row_has_tare, _u8p)
        v.pos_ptr = _abi.ptr(self.pos_ptr, _u64p)
        v.pos_feature = _abi.ptr(self.pos_feature, _u32p)
        v.pos_value = _abi.ptr(self.pos_value, _u32p)
        v.neg_ptr = _abi.ptr(self.neg_ptr, _u64p)
        v.neg_feature = _abi.ptr(self.neg_feature, _u32p)
        return v

    def upload(self, engine, tare_observed=None, tare_values=None):
        """Engine.set_rows_diff: the device (or, in tests, the oracle) expands the diff"""
        F = engine.model.n_features
        to = np.zeros(F, np.uint8) if tare_observed is None else tare_observed
        tv = np.zeros(F, np.uint32) if tare_values is None else tare_values
        engine.set_rows_diff(self.n_rows, to, tv, self.row_has_tare, self.pos_ptr, self.pos_feature, self.pos_value,
                             self.neg_ptr, self.neg_feature)

    def to_columns(self, model_file, tare_observed=None, tare_values=None):
        """dense columnar block on the host (cat8, wide, mask) -- lb_set_rows layout"""
        v = model_file.view
        F = v.n_features
        F8 = sum(1 for i in range(F) if v.specs[i].type in (_abi.LB_BB, _abi.LB_DD))
        R, W = self.n_rows, (self.n_rows + 31) // 32
        cat8 = np.zeros((F8, R), np.uint8)
        wide = np.zeros((F - F8, R), np.uint32)
        mask = np.zeros((F, W), np.uint32)
        to = None if tare_observed is None else np.ascontiguousarray(tare_observed, np.uint8)
        tv = None if tare_values is None else np.ascontiguousarray(tare_values, np.uint32)
        view = self._view()
        _check(lib().lbio_rows_to_columns(model_file.h, C.byref(view), _abi.ptr(to, _u8p), _abi.ptr(tv, _u32p),
                                          _abi.ptr(cat8, _u8p), _abi.ptr(wide, _u32p), _abi.ptr(mask, _u32p)))
        return cat8, wide, mask


def rows_read(path, model_file, first_row=0, max_rows=0, n_threads=0, open_dictionaries=False, seed=0):
    """open_dictionaries: DPD values the model's dictionary lacks are appended to it (with stick-breaking
    betas drawn from `seed`) instead of being an error -- `model_file` is modified; whole file only."""
    v = RowBlockView()
    if open_dictionaries:
        n_new = C.c_int32()
        _check(lib().lbio_rows_read_open(_path(path), model_file.h, seed, n_threads, C.byref(v), C.byref(n_new)))
        model_file.refresh()
    else:
        _check(lib().lbio_rows_read(_path(path), model_file.h, first_row, max_rows, n_threads, C.byref(v)))
    try:
        R = v.n_rows
        pos_ptr = _arr(v.pos_ptr, R + 1, np.uint64)
        neg_ptr = _arr(v.neg_ptr, R + 1, np.uint64)
        return Rows(_arr(v.rowids, R, np.uint64), _arr(v.row_has_tare, R, np.uint8), pos_ptr,
                    _arr(v.pos_feature, pos_ptr[-1], np.uint32), _arr(v.pos_value, pos_ptr[-1], np.uint32), neg_ptr,
                    _arr(v.neg_feature, neg_ptr[-1], np.uint32), v.stream_rows)
    finally:
        lib().lbio_rows_free(C.byref(v))


def rows_write(path, model_file, rows, tare_values=None):
    """tare_values: the tare row's cell values, written as the values of every neg part (what loom's sparsify writes)"""
    view = rows._view()
    if tare_values is None:
        _check(lib().lbio_rows_write(_path(path), model_file.h, C.byref(view)))
        return
    tv = np.ascontiguousarray(tare_values, np.uint32)
    w = C.c_void_p()
    _check(lib().lbio_rows_writer_open(_path(path), C.byref(w)))
    status = lib().lbio_rows_writer_write(w, model_file.h, C.byref(view), _abi.ptr(tv, _u32p))
    closed = lib().lbio_rows_writer_close(w)
    _check(status)
    _check(closed)


def stream_stats(path):
    n, biggest = C.c_uint64(), C.c_uint32()
    _check(lib().lbio_stream_stats(_path(path), C.byref(n), C.byref(biggest)))
    return n.value, biggest.value


def tares_read(path, model_file, open_dictionaries=False, seed=0):
    """open_dictionaries: a DPD value of the tare row that the dictionary lacks is appended (lbio_tares_read_open;
    call after the rows were read with open dictionaries -- the model file is modified)"""
    F = model_file.view.n_features
    to, tv, n = np.zeros(F, np.uint8), np.zeros(F, np.uint32), C.c_int32()
    if open_dictionaries:
        n_new = C.c_int32()
        _check(lib().lbio_tares_read_open(_path(path), model_file.h, seed, _abi.ptr(to, _u8p), _abi.ptr(tv, _u32p),
                                          C.byref(n), C.byref(n_new)))
        model_file.refresh()
    else:
        _check(lib().lbio_tares_read(_path(path), model_file.h, _abi.ptr(to, _u8p), _abi.ptr(tv, _u32p), C.byref(n)))
    return to, tv, n.value


def tares_write(path, model_file, tare_observed, tare_values):
    to = np.ascontiguousarray(tare_observed, np.uint8)
    tv = np.ascontiguousarray(tare_values, np.uint32)
    _check(lib().lbio_tares_write(_path(path), model_file.h, _abi.ptr(to, _u8p), _abi.ptr(tv, _u32p)))


# ------------------------------------------------------------------------------------------ assignments / groups
def assign_read(path):
    b = AssignBlock()
    _check(lib().lbio_assign_read(_path(path), C.byref(b)))
    try:
        return (_arr(b.rowids, b.n_rows, np.uint64),
                _arr(b.groupids, b.n_rows * b.n_kinds, np.uint32).reshape(int(b.n_rows), int(b.n_kinds)))
    finally:
        lib().lbio_assign_free(C.byref(b))


def assign_write(path, rowids, groupids):
    rowids = np.ascontiguousarray(rowids, np.uint64)
    groupids = np.ascontiguousarray(groupids, np.uint32).reshape(len(rowids), -1)
    _check(lib().lbio_assign_write(_path(path), len(rowids), groupids.shape[1], _abi.ptr(rowids, _u64p),
                                   _abi.ptr(groupids, _u32p)))


def groups_read(path, model_file, featureids):
    fids = np.ascontiguousarray(featureids, np.uint32)
    b = GroupBlock()
    _check(lib().lbio_groups_read(_path(path), model_file.h, _abi.ptr(fids, _u32p), len(fids), C.byref(b)))
    try:
        G = b.n_groups
        return (_arr(b.counts, G, np.uint32), _arr(b.ints, b.n_int_rows * G, np.uint32).reshape(b.n_int_rows, G),
                _arr(b.flts, b.n_flt_rows * G, np.float64).reshape(b.n_flt_rows, G))
    finally:
        lib().lbio_groups_free(C.byref(b))


def groups_write(path, model_file, featureids, counts, ints, flts, order=None):
    fids = np.ascontiguousarray(featureids, np.uint32)
    counts = np.ascontiguousarray(counts, np.uint32)
    ints = np.ascontiguousarray(ints, np.uint32)
    flts = np.ascontiguousarray(flts, np.float64)
    order = np.arange(len(counts), dtype=np.uint32) if order is None else np.ascontiguousarray(order, np.uint32)
    _check(lib().lbio_groups_write(_path(path), model_file.h, _abi.ptr(fids, _u32p), len(fids), len(counts),
                                   _abi.ptr(counts, _u32p), _abi.ptr(ints, _u32p), _abi.ptr(flts, _f64p),
                                   _abi.ptr(order, _u32p), len(order)))


def mixture_path(groups_dir, kindid):
    """store::get_mixture_path (src/store.hpp:59-66)"""
    return os.path.join(groups_dir, "mixture.%d.pbs.gz" % kindid)


def sorted_groupids(counts):
    """CrossCat::get_sorted_groupids for one kind (src/cross_cat.cc:212-236): packed ids of the non-empty groups, largest
    first, tied groups in the order the reference's own std::sort call leaves them"""
    counts = np.ascontiguousarray(counts, np.uint32)
    order = np.zeros(max(len(counts), 1), np.uint32)
    n = C.c_int32()
    _check(lib().lbio_sorted_groupids(_abi.ptr(counts, _u32p), len(counts), _abi.ptr(order, _u32p), C.byref(n)))
    return order[:n.value].copy()


def shuffle_stream(messages_in, shuffled_out, seed=0, target_mem_bytes=4e9):
    """loom.runner.shuffle (loom/runner.py:178-194) without the process"""
    _check(lib().lbio_shuffle_stream(_path(messages_in), _path(shuffled_out), seed, float(target_mem_bytes)))
