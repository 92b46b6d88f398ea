"""TEST INFRASTRUCTURE: the sharding logic of loom_b200/dist.py (row-sharded proposer, row-sharded accumulate,
task-sharded hyper kernel) restated over torch.distributed so that it can run at world_size 2 on the CPU box with the
gloo backend and the oracle as the compute backend.  The product path is loom_b200/dist.py: NCCL calls issued inside
libloomb200 (include/loom_b200_comm.h), no torch."""
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


# ---------------------------------------------------------------------------------------------
# Row-sharded bulk accumulate (K2): every rank adds ITS rows to zeroed group tables, the tables
# are combined across ranks, every rank ends with the statistics of all rows.  Integer statistics (clustering counts,
# BB heads / tails, DD / DPD counts, GP count / sum, NICH count) are sums, so one all-reduce(sum) is exact under any
# order.  Float moments: GP log_prod is a sum; NICH (mean, count_times_variance) are merged with the pairwise update of
# Chan et al. over two more all-reduces (sum of n*mean, then sum of ctv + n*(mean - global mean)^2), float64.
def _feature_rows(model, kind):
    """[(type, first int row, int rows, first flt row)] of the features of `kind` in its tables (include/loom_b200.h)"""
    out, i, f = [], 0, 0
    for fid in model.kind_features(kind):
        feat = model.features[fid]
        t = feat["type"]
        ni = {"bb": 2, "gp": 2, "nich": 1}.get(t)
        if ni is None:
            ni = len(feat["alphas"] if t == "dd" else feat["betas"]) + 1
        nf = {"gp": 1, "nich": 2}.get(t, 0)
        out.append((t, i, ni, f))
        i, f = i + ni, f + nf
    return out


def _allreduce_numpy(x, device, group=None):
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(x))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def sharded_accumulate_rows(engine, rank, world, group=None):
    """accumulate_rows of the whole block with rows [rank*R/world, (rank+1)*R/world) added on this rank.
    Every kind must hold groups 0..n-1 (all used by some row) and the assignments of all rows; on return every rank
    holds the full statistics, as a single-rank accumulate_rows would leave them."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if engine.abi.prefix == "lb_" else None
    R = engine.n_rows
    lo, hi = (R * rank) // world, (R * (rank + 1)) // world
    model = engine.model
    for k in range(engine.n_kinds()):
        assign = engine.get_assignments(k)
        n = int(assign[assign != 0xFFFFFFFF].max()) + 1 if (assign != 0xFFFFFFFF).any() else 0
        engine.set_groups(k, n)                       # zeroed tables (also clears the assignments)
        engine.set_assignments(k, assign)
        engine.accumulate_rows(k, lo, hi, +1)
        counts, ints, flts = engine.get_groups(k)
        counts, ints, flts = counts[:n].astype(np.int64), ints[:, :n].astype(np.int64), flts[:, :n].copy()
        if world > 1:
            local_ints = ints.copy()
            counts = _allreduce_numpy(counts, device, group)
            ints = _allreduce_numpy(ints, device, group)
            for t, i0, ni, f0 in _feature_rows(model, k):
                if t == "gp":
                    flts[f0] = _allreduce_numpy(flts[f0], device, group)
                elif t == "nich":
                    n_r, n_all = local_ints[i0].astype(np.float64), ints[i0].astype(np.float64)
                    mean_r, ctv_r = flts[f0].copy(), flts[f0 + 1].copy()
                    mean = _allreduce_numpy(n_r * mean_r, device, group) / np.maximum(n_all, 1.0)
                    flts[f0] = np.where(n_all > 0, mean, 0.0)
                    flts[f0 + 1] = _allreduce_numpy(ctv_r + n_r * (mean_r - mean) ** 2, device, group)
        engine.set_groups(k, n, counts.astype(np.uint32), ints.astype(np.uint32), flts)
        engine.set_assignments(k, assign)


# ---------------------------------------------------------------------------------------------
# Feature-sharded hyper kernel: the tasks of HyperKernel::run are independent
# (src/hyper_kernel.cc:205-234), so rank r runs the tasks t with t % world == r on its replica of the state and the
# chosen values are exchanged: every rank zeroes what it did not choose and one all-reduce(sum) per array puts the
# owner's value everywhere (x + 0 is exact).  Draws are keyed by (seed, task id): the result equals hyper_run's.
def sharded_hyper_run(engine, rank, world, seed, group=None):
    import torch
    K, F = engine.n_kinds(), engine.model.n_features
    owner = np.arange(1 + K + F) % world
    engine.hyper_run(seed, task_mask=(owner == rank).astype(np.uint8))
    if world == 1:
        return
    device = torch.device("cuda", torch.cuda.current_device()) if engine.abi.prefix == "lb_" else None
    specs, aux, kind_py, topo = engine.get_hypers()
    hp = np.array([[specs[f].hp[c] for c in range(4)] for f in range(F)], np.float32)
    aux = np.array(aux, np.float32)
    mine_f = owner[1 + K:] == rank
    hp[~mine_f] = 0
    for f in range(F):
        if not mine_f[f] and specs[f].dim > 0 and specs[f].type in (1, 2):      # DD alphas / DPD betas of unowned features
            aux[specs[f].aux_off:specs[f].aux_off + specs[f].dim] = 0
    kind_py = np.array(kind_py, np.float32)
    kind_py[owner[1:1 + K] != rank] = 0
    topo = np.array(topo, np.float32) * (owner[0] == rank)
    hp = _allreduce_numpy(hp, device, group)
    if len(aux):
        aux = _allreduce_numpy(aux, device, group)
    kind_py = _allreduce_numpy(kind_py, device, group)
    topo = _allreduce_numpy(topo, device, group)
    for f in range(F):
        for c in range(4):
            specs[f].hp[c] = hp[f, c]
    engine.set_hypers(specs, aux, kind_py, topo)
