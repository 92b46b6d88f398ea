"""ctypes binding of the C ABI declared in include/loom_b200.h.

`Abi(path, prefix)` binds one shared library exporting that ABI.  The product
binds `loom_b200/csrc/libloomb200.so` with prefix ``lb_`` (see `load_cuda`);
there is no CPU fallback: if the CUDA library is missing, `load_cuda` raises.
The tests bind the CPU oracle (`oracle/libloom_oracle.so`, prefix ``lo_``)
through the same class so both sides are driven by identical calls.
"""
import ctypes as C
import os

import numpy as np

LB_BB, LB_DD, LB_DPD, LB_GP, LB_NICH = range(5)
TYPE_NAMES = {"bb": LB_BB, "dd": LB_DD, "dpd": LB_DPD, "gp": LB_GP, "nich": LB_NICH}
UNASSIGNED = 0xFFFFFFFF
DPD_OTHER = 0xFFFFFFFF   # LB_DPD_OTHER: query_sample drew a value outside the dictionary


class FeatureSpec(C.Structure):
    _fields_ = [("type", C.c_int32), ("dim", C.c_int32), ("hp", C.c_float * 4),
                ("aux_off", C.c_int32)]


class KindInfo(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("n_groups", C.c_int32),
                ("n_int_rows", C.c_int32), ("n_flt_rows", C.c_int32),
                ("sample_size", C.c_uint64), ("py_alpha", C.c_float), ("py_d", C.c_float)]


_HP_FIELDS = ["topology", "clustering", "bb_alpha", "bb_beta", "dd_alpha", "dpd_gamma",
              "dpd_alpha", "gp_alpha", "gp_inv_beta", "nich_mu", "nich_kappa",
              "nich_sigmasq", "nich_nu"]


class HyperPrior(C.Structure):
    _fields_ = sum(([(n, C.POINTER(C.c_float)), ("n_" + n, C.c_int32)] for n in _HP_FIELDS), [])


_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_h = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/loom_b200.h
SIGNATURES = {
    "last_error": (C.c_char_p, []),
    "create": (C.c_int, [C.c_int, C.POINTER(_h)]),
    "destroy": (None, [_h]),
    "set_model": (C.c_int, [_h, C.c_int32, C.POINTER(FeatureSpec), _f32p, C.c_int32, C.c_int32,
                            _u32p, _f32p, _f32p, C.c_int32]),
    "set_rows": (C.c_int, [_h, C.c_uint64, _u8p, _u32p, _u32p]),
    "reload_rows": (C.c_int, [_h, C.c_uint64, _u8p, _u32p, _u32p]),
    "set_rows_diff": (C.c_int, [_h, C.c_uint64, _u8p, _u32p, _u8p, _u64p, _u32p, _u32p, _u64p, _u32p]),
    "query_score": (C.c_int, [_h, C.c_uint64, C.c_uint64, _f32p]),
    "query_sample": (C.c_int, [_h, C.c_uint64, _u32p, C.c_int32, C.c_uint32, C.c_uint64, C.c_uint64, _u32p, _u32p]),
    "share_rows": (C.c_int, [_h, _h]),
    "generate_rows": (C.c_int, [_h, C.c_uint64, C.c_float, C.c_uint64]),
    "get_rows": (C.c_int, [_h, _u8p, _u32p, _u32p]),
    "differ_reset": (C.c_int, [_h]),
    "differ_add_rows": (C.c_int, [_h, C.POINTER(C.c_int32), _u32p]),
    "differ_make_tare": (C.c_int, [_h, C.POINTER(C.c_int32), _u32p, _u8p, _u32p]),
    "differ_compress_rows": (C.c_int, [_h, _u8p, _u32p, _u64p, _u64p]),
    "differ_fetch": (C.c_int, [_h, _u64p, _u32p, _u32p, _u64p, _u32p]),
    "cat_sweep_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, _u32p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]),
    "cat_sweep_streams": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, _u32p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]),
    "cat_sweep": (C.c_int, [_h, _u32p, C.c_uint64, C.c_uint64, C.c_uint64, _u8p, _u32p, _f32p,
                            C.c_int32]),
    "get_assignments": (C.c_int, [_h, C.c_int32, _u32p]),
    "set_assignments": (C.c_int, [_h, C.c_int32, _u32p]),
    "kind_info": (C.c_int, [_h, C.c_int32, C.POINTER(KindInfo)]),
    "get_groups": (C.c_int, [_h, C.c_int32, _u32p, _u32p, _f64p]),
    "set_groups": (C.c_int, [_h, C.c_int32, C.c_int32, _u32p, _u32p, _f64p]),
    "score_rows": (C.c_int, [_h, C.c_int32, C.c_uint64, C.c_uint64, _f32p]),
    "score_rows_diff": (C.c_int, [_h, C.c_int32, C.c_uint64, C.c_uint64, _f32p]),
    "accumulate_rows": (C.c_int, [_h, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32]),
    "score_data": (C.c_int, [_h, _f64p]),
    "kind_proposer_score": (C.c_int, [_h, _f32p]),
    "proposer_accumulate": (C.c_int, [_h, C.c_uint64, C.c_uint64]),
    "proposer_tables": (C.c_int, [_h, C.c_int32, C.POINTER(C.c_void_p), _u64p, C.POINTER(C.c_void_p), _u64p]),
    "proposer_finish": (C.c_int, [_h, _f32p]),
    "kind_run": (C.c_int, [_h, C.c_int32, C.c_uint64, _u32p, C.POINTER(C.c_int32)]),
    "init_featureless_kinds": (C.c_int, [_h, C.c_int32, C.c_uint64]),
    "prepare_featureless_kinds": (C.c_int, [_h, C.c_int32, C.c_uint64]),
    "get_kinds": (C.c_int, [_h, C.POINTER(C.c_int32), _u32p]),
    "set_hyper_prior": (C.c_int, [_h, C.POINTER(HyperPrior)]),
    "hyper_run": (C.c_int, [_h, C.c_uint64]),
    "hyper_run_tasks": (C.c_int, [_h, C.c_uint64, _u8p]),
    "set_hypers": (C.c_int, [_h, C.POINTER(FeatureSpec), _f32p, _f32p, _f32p]),
    "get_hypers": (C.c_int, [_h, C.POINTER(FeatureSpec), _f32p, _f32p, _f32p]),
    "sync": (C.c_int, [_h]),
    "timer_start": (C.c_int, [_h]),
    "timer_stop": (C.c_int, [_h, _f32p]),
    "launch_count": (C.c_int, [_h, _u64p, C.c_int32, C.c_int32]),
    "kernel_ms": (C.c_int, [_h, _f64p, C.c_int32, C.c_int32]),
}


class CommRoundStats(C.Structure):
    """lb_comm_round_stats (include/loom_b200_comm.h)"""
    _fields_ = [(n, C.c_double) for n in ("accumulate_ms", "allreduce_ms", "allreduce_bytes", "score_ms", "allgather_ms",
                                          "allgather_bytes", "row_begin", "row_end", "feature_begin", "feature_end")]


# include/loom_b200_comm.h: NCCL entry points, CUDA library only (the oracle has no collectives)
COMM_SIGNATURES = {
    "comm_unique_id": (C.c_int, [_u8p]),
    "comm_init": (C.c_int, [_h, _u8p, C.c_int32, C.c_int32]),
    "comm_destroy": (C.c_int, [_h]),
    "comm_rank": (C.c_int, [_h, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "comm_barrier": (C.c_int, [_h]),
    "comm_max_f64": (C.c_int, [_h, _f64p, C.c_int32]),
    "comm_kind_proposer_score": (C.c_int, [_h, _f32p, C.POINTER(CommRoundStats)]),
    "comm_hyper_run": (C.c_int, [_h, C.c_uint64, _f64p]),
    "comm_accumulate_rows": (C.c_int, [_h, _f64p, _f64p]),
    "comm_sync_kinds": (C.c_int, [_h, C.POINTER(C.c_int32), _f64p, _f64p]),
}


class LoomError(RuntimeError):
    """Non-zero status from the C ABI (the reference aborts: src/common.hpp:52-59)."""


def ptr(a, ty):
    if a is None:
        return None
    return a.ctypes.data_as(ty)


class Abi:
    def __init__(self, path, prefix):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "lb_" else C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        if prefix == "lb_":
            for name, (res, args) in COMM_SIGNATURES.items():
                fn = getattr(self.lib, prefix + name)
                fn.restype = res
                fn.argtypes = args
                setattr(self, name, fn)

    def check(self, status):
        if status != 0:
            msg = self.last_error()
            raise LoomError((msg or b"?").decode(errors="replace"))


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_LIB = os.path.join(_REPO, "loom_b200", "csrc", "libloomb200.so")
_cuda = None


def load_cuda():
    """Bind the CUDA library.  Fails loudly if it has not been built."""
    global _cuda
    if _cuda is None:
        override = os.environ.get("LOOM_B200_LIB")      # a build variant of the same library (csrc/Makefile `variants`)
        if override:
            if not os.path.exists(override):
                raise LoomError("LOOM_B200_LIB=%s does not exist" % override)
            _cuda = Abi(override, "lb_")
            return _cuda
        if not os.path.exists(CUDA_LIB):
            raise LoomError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % CUDA_LIB)
        _cuda = Abi(CUDA_LIB, "lb_")
    return _cuda
