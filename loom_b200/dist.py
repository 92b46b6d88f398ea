"""Multi-GPU plumbing for the one path that has a real exchange step: the row-sharded kind
proposer.  Every rank accumulates its own row shard into the all-feature
proposer tables, the tables are summed across ranks (NCCL over NVLink for the CUDA library;
gloo for the CPU oracle in tests), and every rank then scores the identical full tables.

torch.distributed is used for the collective only; the tables stay where the engine put them
(device pointers are wrapped zero-copy through __cuda_array_interface__).
"""
import ctypes as C

import numpy as np


class _DeviceArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _wrap(engine, ptr, n, kind):
    import torch
    if n == 0:
        return None
    if engine.abi.prefix == "lb_":   # device memory owned by libloomb200
        typestr = "<i4" if kind == "int" else "<f8"
        return torch.as_tensor(_DeviceArray(ptr, n, typestr), device="cuda")
    ctype = C.c_int32 if kind == "int" else C.c_double   # host memory owned by the oracle
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(int(n),))
    return torch.from_numpy(arr)


def allreduce_proposer_tables(engine, group=None):
    """Sum the (partial) proposer tables of every kind over all ranks, in place."""
    import torch.distributed as dist
    engine.sync()
    works = []
    for k in range(engine.n_kinds()):
        pi, ni, pf, nf = engine.proposer_tables(k)
        for t in (_wrap(engine, pi, ni, "int"), _wrap(engine, pf, nf, "flt")):
            if t is not None:   # uint32 counts are summed as int32: two's-complement addition is the same
                works.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True), t))
    for w, _ in works:
        w.wait()
    if engine.abi.prefix == "lb_":
        import torch
        torch.cuda.synchronize()


def sharded_kind_proposer_score(engine, rank, world, fetch=True, group=None):
    """kind_proposer_score with rows [rank*R/world, (rank+1)*R/world) accumulated on this rank."""
    R = engine.n_rows
    lo, hi = (R * rank) // world, (R * (rank + 1)) // world
    engine.proposer_accumulate(lo, hi)
    if world > 1:
        allreduce_proposer_tables(engine, group)
    return engine.proposer_finish(fetch=fetch)
