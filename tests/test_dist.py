"""N > 1 path.
CPU: world_size 2 over gloo with the oracle as compute backend -- the sharding logic (which rows / features / tasks a
     rank takes, what is summed, what must come out) restated over torch.distributed in tests/dist_gloo.py.
GPU: world_size 2 over NCCL through the library's own entry points (include/loom_b200_comm.h, loom_b200/dist.py):
     two plain processes, no torch; skipped with fewer than 2 devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, os.path.join({repo!r}, "tests"))
import numpy as np
import torch, torch.distributed as dist
from loom_b200 import Engine, _abi, synth, sweep_actions
from dist_gloo import sharded_kind_proposer_score
backend = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if backend == "nccl":
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    abi = _abi.load_cuda()
else:
    dist.init_process_group("gloo")
    abi = _abi.Abi(os.path.join({repo!r}, "oracle", "libloom_oracle.so"), "lo_")
model, rows, truth = synth.generate(3000, [("bb", 3), ("dd16", 3), ("dd256", 1), ("gp", 3), ("nich", 3)], density=0.6, seed=13)
e = Engine(abi, device=rank if backend == "nccl" else 0)
e.set_model(model); e.set_rows(rows)
for k in range(model.n_kinds):      # identical state on every rank
    a = truth["assign"][k].astype(np.uint32)
    uniq, inv = np.unique(a, return_inverse=True)
    e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
e.accumulate_rows()
e.init_featureless_kinds(4, seed=3)
full = e.kind_proposer_score()                      # single-rank reference
sharded = sharded_kind_proposer_score(e, rank, world)
err = np.abs(sharded.astype(np.float64) - full) / np.maximum(1.0, np.abs(full))
assert err.max() <= 1e-6, err.max()
f2k, n = e.kind_run(8, seed=5)                       # same seed -> same structure on every rank
out = torch.tensor(f2k.astype(np.int64))
if backend == "nccl": out = out.cuda()
gathered = [torch.zeros_like(out) for _ in range(world)]
dist.all_gather(gathered, out)
assert all(torch.equal(g, gathered[0]) for g in gathered)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok.%d" % rank), "w").write("%g %d" % (float(err.max()), int(n)))
dist.destroy_process_group()
'''


KIND_SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r})
import numpy as np
import torch, torch.distributed as dist
from loom_b200 import Engine, _abi, synth, sweep_actions
backend = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if backend == "nccl":
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    abi = _abi.load_cuda()
else:
    dist.init_process_group("gloo")
    abi = _abi.Abi(os.path.join({repo!r}, "oracle", "libloom_oracle.so"), "lo_")
model, rows, _ = synth.generate(800, [("bb", 4), ("dd16", 4), ("gp", 4), ("nich", 3)], density=0.6, seed=17)
K = model.n_kinds
acts = np.concatenate([sweep_actions(800), sweep_actions(800, add=True, remove=True)])
# kinds are independent (doc/adapting.md:144-147): rank r sweeps kinds k with k % world == r
e = Engine(abi, device=rank if backend == "nccl" else 0)
e.set_model(model); e.set_rows(rows)
mask = np.array([k % world == rank for k in range(K)], np.uint8)
e.cat_sweep(acts, seed=21, kind_mask=mask)
mine = np.stack([e.get_assignments(k) if mask[k] else np.zeros(800, np.uint32) for k in range(K)]).astype(np.int64)
t = torch.tensor(mine)
if backend == "nccl": t = t.cuda()
dist.all_reduce(t)                                   # gather of disjoint kind columns
ref = Engine(abi, device=rank if backend == "nccl" else 0)
ref.set_model(model); ref.set_rows(rows)
ref.cat_sweep(acts, seed=21)                         # all kinds on one device
full = np.stack([ref.get_assignments(k) for k in range(K)]).astype(np.int64)
assert np.array_equal(t.cpu().numpy(), full)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok.%d" % rank), "w").write("1")
dist.destroy_process_group()
'''


ACCUMULATE_WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, os.path.join({repo!r}, "tests"))
import numpy as np
import torch, torch.distributed as dist
from loom_b200 import Engine, _abi, synth
from dist_gloo import sharded_accumulate_rows
backend = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if backend == "nccl":
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    abi = _abi.load_cuda()
else:
    dist.init_process_group("gloo")
    abi = _abi.Abi(os.path.join({repo!r}, "oracle", "libloom_oracle.so"), "lo_")
model, rows, truth = synth.generate(3001, [("bb", 3), ("dd16", 2), ("dpd5", 2), ("gp", 3), ("nich", 3)], density=0.6, seed=19)
def engine():
    e = Engine(abi, device=rank if backend == "nccl" else 0)
    e.set_model(model); e.set_rows(rows)
    for k in range(model.n_kinds):
        uniq, inv = np.unique(truth["assign"][k].astype(np.uint32), return_inverse=True)
        e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
    return e
ref = engine(); ref.accumulate_rows()                 # all rows on one rank
e = engine(); sharded_accumulate_rows(e, rank, world)
for k in range(model.n_kinds):
    (c0, i0, f0), (c1, i1, f1) = ref.get_groups(k), e.get_groups(k)
    assert np.array_equal(c0, c1) and np.array_equal(i0, i1)                       # integers: bit-exact
    assert np.allclose(f0, f1, rtol=1e-9, atol=1e-9)                               # float64 moments, merged pairwise
    assert np.array_equal(ref.get_assignments(k), e.get_assignments(k))
    assert np.abs(ref.score_rows(k, 0, 64) - e.score_rows(k, 0, 64)).max() <= 1e-5
assert abs(ref.score_data() - e.score_data()) <= 1e-6 * abs(ref.score_data())
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok.%d" % rank), "w").write("1")
dist.destroy_process_group()
'''


HYPER_WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, os.path.join({repo!r}, "tests"))
import numpy as np
import torch, torch.distributed as dist
from loom_b200 import Engine, _abi, synth
from dist_gloo import sharded_hyper_run
from test_dpd import GRIDS
backend = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if backend == "nccl":
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    abi = _abi.load_cuda()
else:
    dist.init_process_group("gloo")
    abi = _abi.Abi(os.path.join({repo!r}, "oracle", "libloom_oracle.so"), "lo_")
model, rows, truth = synth.generate(1500, [("bb", 3), ("dd16", 2), ("dpd5", 2), ("gp", 3), ("nich", 3)], density=0.6, seed=23)
grids = dict(GRIDS, bb={{"alpha": [0.5, 2.0], "beta": [0.5, 2.0]}})
def engine():
    e = Engine(abi, device=rank if backend == "nccl" else 0)
    e.set_model(model); e.set_rows(rows)
    for k in range(model.n_kinds):
        uniq, inv = np.unique(truth["assign"][k].astype(np.uint32), return_inverse=True)
        e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
    e.accumulate_rows()
    e.set_hyper_prior(grids)
    return e
ref, e = engine(), engine()
for seed in (1, 2):                                    # two rounds: the second starts from exchanged hypers
    ref.hyper_run(seed)
    sharded_hyper_run(e, rank, world, seed)
    (s0, a0, k0, t0), (s1, a1, k1, t1) = ref.get_hypers(), e.get_hypers()
    for f in range(model.n_features):
        assert [s0[f].hp[c] for c in range(4)] == [s1[f].hp[c] for c in range(4)], f
    assert np.array_equal(a0, a1) and np.array_equal(k0, k1) and np.array_equal(t0, t1)
    assert ref.score_data() == e.score_data()
    for k in range(model.n_kinds):
        assert np.array_equal(ref.score_rows(k, 0, 32), e.score_rows(k, 0, 32))
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok.%d" % rank), "w").write("1")
dist.destroy_process_group()
'''


NATIVE_WORKER = r'''
# two plain processes, NCCL through libloomb200's own entry points (no torch anywhere)
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, os.path.join({repo!r}, "tests"))
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions, dist
from test_dpd import GRIDS
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
abi = _abi.load_cuda()
model, rows, truth = synth.generate(3000, [("bb", 3), ("dd16", 3), ("dd256", 1), ("dpd5", 2), ("gp", 3), ("nich", 3)], density=0.6, seed=13)
K = model.n_kinds
def engine(assigned=True):
    e = Engine(abi, device=rank)
    e.set_model(model); e.set_rows(rows)
    if assigned:
        for k in range(K):      # identical state on every rank
            uniq, inv = np.unique(truth["assign"][k].astype(np.uint32), return_inverse=True)
            e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
        e.accumulate_rows()
    return e
e = engine()
dist.init(e, rank, world, tag={tag!r})
# --- row-sharded proposer: K4 on this rank's rows, all-reduce, K5 on this rank's features, all-gather
e.init_featureless_kinds(4, seed=3)
full = e.kind_proposer_score()                      # single-rank reference on the same device
sharded, stats = dist.sharded_kind_proposer_score(e)
err = np.abs(sharded.astype(np.float64) - full) / np.maximum(1.0, np.abs(full))
assert err.max() <= 1e-6, err.max()
assert stats["allreduce_bytes"] > 0 and stats["allgather_bytes"] > 0 and stats["row_end"] - stats["row_begin"] == 1500
f2k, n = e.kind_run(8, seed=5)                       # same seed, same scores -> same structure on every rank
chk = dist.max_over_ranks(e, np.concatenate([f2k, -f2k.astype(np.float64)]))
assert np.array_equal(chk[:len(f2k)], f2k) and np.array_equal(-chk[len(f2k):], f2k)    # max == min == mine
# --- task-sharded hyper kernel == hyper_run, bit for bit
grids = dict(GRIDS, bb={{"alpha": [0.5, 2.0], "beta": [0.5, 2.0]}})
ref, h = engine(), engine()
dist.init(h, rank, world, tag={tag!r} + "h")
ref.set_hyper_prior(grids); h.set_hyper_prior(grids)
for seed in (1, 2):
    ref.hyper_run(seed)
    dist.sharded_hyper_run(h, seed)
    (s0, a0, k0, t0), (s1, a1, k1, t1) = ref.get_hypers(), h.get_hypers()
    for f in range(model.n_features):
        assert [s0[f].hp[c] for c in range(4)] == [s1[f].hp[c] for c in range(4)], f
    assert np.array_equal(a0, a1) and np.array_equal(k0, k1) and np.array_equal(t0, t1)
    assert abs(ref.score_data() - h.score_data()) <= 1e-12 * abs(ref.score_data())
# --- kind-sharded sweep + publication of the kinds == the sweep of all kinds on one device
acts = np.concatenate([sweep_actions(3000), sweep_actions(3000, add=True, remove=True)])
one, sh = engine(False), engine(False)
dist.init(sh, rank, world, tag={tag!r} + "s")
one.cat_sweep(acts, seed=21)
owner = np.arange(K) % world
sh.cat_sweep(acts, seed=21, kind_mask=(owner == rank).astype(np.uint8))
nbytes, ms = dist.sync_kinds(sh, owner)
assert nbytes > 0
for k in range(K):
    assert np.array_equal(one.get_assignments(k), sh.get_assignments(k)), k
    for a, b in zip(one.get_groups(k), sh.get_groups(k)):
        assert np.array_equal(a, b), k
    assert np.array_equal(one.score_rows(k, 0, 64), sh.score_rows(k, 0, 64))
assert abs(one.score_data() - sh.score_data()) <= 1e-12 * abs(one.score_data())
more = sweep_actions(3000, add=True, remove=True)    # and the chain goes on from the published state
one.cat_sweep(more, seed=22, action_offset=len(acts)); sh.cat_sweep(more, seed=22, action_offset=len(acts))
for k in range(K):
    assert np.array_equal(one.get_assignments(k), sh.get_assignments(k)), k
# --- row-sharded bulk accumulate (K2) == the single-device accumulate: integers bit-exact, float64 moments to 1e-9
acc_ref, acc = engine(), engine(False)
for k in range(K):
    uniq, inv = np.unique(truth["assign"][k].astype(np.uint32), return_inverse=True)
    acc.set_groups(k, len(uniq)); acc.set_assignments(k, inv.astype(np.uint32))
dist.init(acc, rank, world, tag={tag!r} + "a")
ar_ms, ar_bytes = dist.sharded_accumulate_rows(acc)
assert ar_bytes > 0
for k in range(K):
    (c0, i0, f0), (c1, i1, f1) = acc_ref.get_groups(k), acc.get_groups(k)
    assert np.array_equal(c0, c1) and np.array_equal(i0, i1), k
    assert np.allclose(f0, f1, rtol=1e-9, atol=1e-9), k
    assert acc_ref.kind_info(k).sample_size == acc.kind_info(k).sample_size
    assert np.abs(acc_ref.score_rows(k, 0, 64) - acc.score_rows(k, 0, 64)).max() <= 1e-5
assert abs(acc_ref.score_data() - acc.score_data()) <= 1e-9 * abs(acc_ref.score_data())
dist.barrier(e)
print("rank %d ok: proposer err %.2e allreduce %.0f B %.3f ms, allgather %.0f B, sync_kinds %.0f B %.3f ms" % (
    rank, err.max(), stats["allreduce_bytes"], stats["allreduce_ms"], stats["allgather_bytes"], nbytes, ms), flush=True)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ok.%d" % rank), "w").write("1")
'''


def _run(backend, tmp_path, worker=None):
    script = tmp_path / "worker.py"
    script.write_text((worker or WORKER).format(repo=REPO))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), backend],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert (tmp_path / "ok.0").exists() and (tmp_path / "ok.1").exists(), out.stdout + out.stderr[-2000:]


def test_row_sharded_proposer_gloo_cpu(oracle_abi, tmp_path):
    _run("gloo", tmp_path)


def test_kind_sharded_sweep_gloo_cpu(oracle_abi, tmp_path):
    """cat sweep with kinds sharded over 2 ranks == the same sweep on one rank (no data-path collective)."""
    _run("gloo", tmp_path, KIND_SHARD_WORKER)


def test_row_sharded_accumulate_gloo_cpu(oracle_abi, tmp_path):
    """bulk accumulate (K2) with rows sharded over 2 ranks and the group tables combined == one rank over all rows"""
    _run("gloo", tmp_path, ACCUMULATE_WORKER)


def test_feature_sharded_hyper_gloo_cpu(oracle_abi, tmp_path):
    """HyperKernel::run with its tasks split over 2 ranks and the chosen hypers exchanged == hyper_run, bit for bit"""
    _run("gloo", tmp_path, HYPER_WORKER)


@pytest.mark.gpu
def test_native_nccl_sharding_two_gpus(cuda_abi, tmp_path):
    """lb_comm_*: row-sharded proposer (all-reduce + all-gather), task-sharded hyper kernel, kind-sharded sweep with
    the kinds published by broadcast -- each against the single-device result, on 2 GPUs, without torch."""
    n_dev = int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(NATIVE_WORKER.format(repo=REPO, tag="t%d" % os.getpid()))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    log_dir = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(log_dir):
        open(os.path.join(log_dir, "native_nccl_test.log"), "w").write("\n".join(outs))
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    assert (tmp_path / "ok.0").exists() and (tmp_path / "ok.1").exists()
