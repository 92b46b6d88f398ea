"""Multi-GPU front end over the NCCL entry points of libloomb200 (include/loom_b200_comm.h).  One process per GPU;
the collectives are issued INSIDE the library on the engine's stream -- no torch, no Python in the data path.

What shards and how:
  kind proposer  rows sharded for the accumulate (K4), ONE ncclAllReduce over the uint32 slab and one over the
                 float64 slab, features sharded for the score phase (K5), ONE ncclAllGather of L[f][k]
                 (src/kind_proposer.cc:323-349 is the OpenMP loop this replaces)
  hyper kernel   task t on rank t % world (src/hyper_kernel.cc:203-234), one all-reduce of the chosen values
  cat sweep      kinds are independent (doc/adapting.md:144-147): lb_cat_sweep(kind_mask), no collective
The only host-side exchange is the 128-byte ncclUniqueId, which travels through a file next to the rendezvous port.
"""
import ctypes as C
import os
import tempfile
import time

import numpy as np

from ._abi import CommRoundStats, ptr, _f32p, _f64p, _u8p

ID_BYTES = 128


def exchange_unique_id(abi, rank, world, tag=None, timeout=120.0):
    """rank 0 calls ncclGetUniqueId and publishes the bytes in a file; the other ranks wait for it."""
    # all ranks of one launch are children of one launcher process (torchrun's agent, a test's Popen loop): its pid makes
    # the file name unique to the launch, so a file a crashed earlier run left behind under the same port is never read
    tag = "%s_%s_%d" % (tag or os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "run"), os.getppid())
    path = os.path.join(tempfile.gettempdir(), "loom_b200_nccl_id_%s" % tag)
    if rank == 0:
        buf = np.zeros(ID_BYTES, np.uint8)
        abi.check(abi.comm_unique_id(ptr(buf, _u8p)))
        tmp = path + ".tmp.%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(buf.tobytes())
        os.replace(tmp, path)          # atomic: readers never see a partial id
        return buf, path
    t0 = time.time()
    while True:
        try:
            data = open(path, "rb").read()
            if len(data) == ID_BYTES:
                return np.frombuffer(data, np.uint8).copy(), path
        except OSError:
            pass
        if time.time() - t0 > timeout:
            raise RuntimeError("rank %d: no NCCL id at %s after %.0f s" % (rank, path, timeout))
        time.sleep(0.02)


def init(engine, rank, world, tag=None):
    """ncclCommInitRank for `engine` (one communicator per engine, on the engine's device)."""
    a = engine.abi
    ident, path = exchange_unique_id(a, rank, world, tag)
    a.check(a.comm_init(engine.h, ptr(ident, _u8p), rank, world))
    barrier(engine)
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass
    return rank, world


def barrier(engine):
    engine.abi.check(engine.abi.comm_barrier(engine.h))


def max_over_ranks(engine, values):
    x = np.ascontiguousarray(np.atleast_1d(values), np.float64).copy()
    engine.abi.check(engine.abi.comm_max_f64(engine.h, ptr(x, _f64p), len(x)))
    return x


def sharded_kind_proposer_score(engine, fetch=True):
    """KindProposer score phase sharded over the communicator's ranks; returns (L[f][k] or None, stats dict)."""
    K = engine.n_kinds()
    out = np.empty((engine.model.n_features, K), np.float32) if fetch else None
    st = CommRoundStats()
    engine.abi.check(engine.abi.comm_kind_proposer_score(engine.h, ptr(out, _f32p), C.byref(st)))
    return out, {n: getattr(st, n) for n, _ in CommRoundStats._fields_}


def sharded_hyper_run(engine, seed):
    """HyperKernel::run with the tasks dealt over the ranks; returns the all-reduce's device time in ms."""
    ms = C.c_double()
    engine.abi.check(engine.abi.comm_hyper_run(engine.h, seed, C.byref(ms)))
    return ms.value


def sync_kinds(engine, owner):
    """publish every kind's state from rank owner[k] after a kind-sharded sweep; returns (payload bytes, device ms)"""
    owner = np.ascontiguousarray(owner, np.int32)
    assert len(owner) == engine.n_kinds()
    b, ms = C.c_double(), C.c_double()
    engine.abi.check(engine.abi.comm_sync_kinds(engine.h, owner.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(b), C.byref(ms)))
    return b.value, ms.value


def sharded_accumulate_rows(engine):
    """bulk accumulate with the rows dealt over the ranks and the tables all-reduced; returns (all-reduce ms, bytes)"""
    ms, b = C.c_double(), C.c_double()
    engine.abi.check(engine.abi.comm_accumulate_rows(engine.h, C.byref(ms), C.byref(b)))
    return ms.value, b.value
