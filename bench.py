#!/usr/bin/env python
"""bench.py -- loom's two headline rates on BASELINE.json configs[2] (C3): cat-kernel cells/sec and kind-kernel
feature-scores/sec, full inference with the kind kernel.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload "C3"): synthetic 1M rows x 256 Dirichlet-Discrete(256) features, density 0.5, 9 kinds from
loom's generator + 32 ephemeral kinds (loom/config.py:55-61), `samples` = 10 independent chains (loom's default
sample_count, loom/tasks.py:49) over ONE shared row block -- the largest single-GPU configuration of BASELINE.json
(C4 / C5 need the windowed row stream; C2 is `--workload C2`).

One STEP is one steady-state batch of Loom::infer_multi_pass with kind structure (src/loom.cc:256-313), for every
chain:
    cat kernel   REMOVE + ADD of every row in stream order (one extra pass; CatKernel via KindKernel::add_row /
                 remove_row, all K + 32 kinds)                                       2 R row actions per chain
    kind kernel  KindKernel::try_run: proposer accumulate (K4) + F x (K+E) marginal scores (K5) + 32 block
                 Pitman-Yor iterations + move_features + fresh ephemeral kinds (src/kind_kernel.cc:118-157)
(HyperKernel::run is `--hyper`: BASELINE.json names hyper-grid inference for C5 only; with loom's default grids a
DD-256 column costs 256 coordinates x 12 grid points of full marginals and would make the step a hyper-kernel benchmark
-- it is timed on its own in roofline_aux.hyper_kernel*.)
A "cell" is one observed (row, feature) value consumed by one cat-kernel action: a step consumes
2 x (#observed cells) per chain.  `value` = cells/s over the WHOLE step (the kind kernel's time included),
inputs resident in HBM, CUDA events.  `kind_kernel` carries the second half of the metric.

e2e     = the same step from HOST buffers: the row block is streamed in again from pinned memory every step
          (lb_reload_rows: the reference re-reads its rows file on every pass, src/stream_interval.hpp) and the
          assignments of every feature-bearing kind of every chain are read back, all inside the timed region.
roofline= the dominant kernel (k_cat_sweep_spec): algorithmic bytes per launch (cells + mask + row id
          + 4 B / kind read and written per row action) / its CUDA-event time against the measured HBM bandwidth.
          Exact sequential Gibbs is latency-bound (a dependent chain per (chain, kind) CTA); roofline_aux carries
          the bandwidth-shaped kernels (K1 score_rows, K2 accumulate, K4, K5) on the same workload.
N > 1   : chains are independent (loom runs one process per sample) -> every rank sweeps its own `samples` chains,
          "scaling": "weak", no data-path collective in the headline.  The path WITH an exchange step is the kind
          kernel of ONE chain: `kind_kernel.sharded` times one proposal round of an identical chain on all N ranks --
          rows sharded for K4, ONE ncclAllReduce per slab, features sharded for K5, ONE ncclAllGather -- plus the
          task-sharded hyper kernel and the kind-sharded sweep's publication step (lb_comm_*, NCCL called from inside
          libloomb200; no torch anywhere in this file).  Its rate is strong-scaled: one chain, N GPUs.
--impl reference times the restated reference (oracle/, float64 C, the faster of the exact-libm and -ffast-math
builds) on the host cores: the same step on a bounded row sample (cpu_baseline.sample).
"""
import argparse
import concurrent.futures
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# one hardware work queue per chain stream (the default of 8 would falsely serialise 21 streams)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "cat_kernel_cells_per_sec"
UNIT = "cells/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload(name, rows, with_truth=False):
    from loom_b200 import synth
    from loom_b200.engine import Model, RowBlock
    cfg = dict(synth.CONFIGS[name])
    if rows:
        cfg["rows"] = rows
    cache = os.path.join(tempfile.gettempdir(), "loom_b200_%s_%d_%d_v2.npz" % (name, cfg["rows"], cfg["seed"]))
    feats = synth.make_features(cfg["types"])
    if os.path.exists(cache):
        z = np.load(cache)
        model = Model(feats, z["f2k"], kind_py=[synth.CLUSTERING] * (int(z["f2k"].max()) + 1),
                      topology=synth.CLUSTERING)
        out = (model, RowBlock(cfg["rows"], z["cat8"], z["wide"], z["mask"]))
        return out + (z["truth"],) if with_truth else out
    model, rows_, truth = synth.generate(**cfg)
    truth = np.stack(truth["assign"]).astype(np.uint32)        # [K][R] generating partition
    try:
        tmp = cache + ".%d.tmp.npz" % os.getpid()                 # ranks of one box race for the cache: write, then rename
        np.savez(tmp, f2k=model.f2k, cat8=rows_.cat8, wide=rows_.wide, mask=rows_.mask, truth=truth)
        os.replace(tmp, cache)
    except OSError:
        pass
    return (model, rows_, truth) if with_truth else (model, rows_)


def step_actions(R):
    """C2-style self-contained pass (benchmarks/sweep_bench.py): ADD all + REMOVE/ADD all, then REMOVE all."""
    from loom_b200 import sweep_actions
    first = np.concatenate([sweep_actions(R), sweep_actions(R, add=True, remove=True)])
    drain = sweep_actions(R, add=False)
    return first, drain


def install_truth(eng, truth):
    """the generating partition as the chain's state (tests do the same): identical on every rank, no Gibbs needed"""
    for k in range(truth.shape[0]):
        uniq, inv = np.unique(truth[k], return_inverse=True)
        eng.set_groups(k, len(uniq))
        eng.set_assignments(k, inv.astype(np.uint32))
    eng.accumulate_rows()


def algorithmic_bytes_per_action(model, rows, n_kinds=None):
    """per row-action, input cells (BB 1 B, DD 1 B, DPD/GP/NICH 4 B
    per observed cell) + 1 bit/feature mask + 8 B row id, + 4 B/kind read and
    4 B/kind written (assignment ids; n_kinds counts the ephemeral kinds when they are swept too)."""
    width = np.array([1 if f["type"] in ("bb", "dd") else 4 for f in model.features], np.float64)
    per_feature = rows.cells_per_feature().astype(np.float64)
    cell_bytes = float((per_feature * width).sum()) / rows.n_rows
    return cell_bytes + model.n_features / 8.0 + 8.0 + 8.0 * (n_kinds or model.n_kinds)


def measure_ingest(abi, model, rows, device=0):
    """Row ingest: the workload's rows as loom's `rows.pbs.gz` (a gzip stream of
    protobuf Row messages), decoded by the native codec (gunzip + framing serial, message decode on all
    host threads) and expanded on the device by lb_set_rows_diff.  Replaces the unzip / parse / split
    stages of src/cat_pipeline.cc:63-87.  Host wall clock; the file is written outside the timed region."""
    from loom_b200 import pbio
    from loom_b200.engine import Engine
    mf = pbio.ModelFile.create(model)
    fd, path = tempfile.mkstemp(suffix=".rows.pbs.gz")
    os.close(fd)
    try:
        pbio.rows_write(path, mf, pbio.Rows.from_block(rows))
        file_bytes = os.path.getsize(path)
        n_msgs, _ = pbio.stream_stats(path)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            view = pbio.RowBlockView()
            pbio._check(pbio.lib().lbio_rows_read(os.fsencode(path), mf.h, 0, 0, 0, ctypes.byref(view)))
            t1 = time.perf_counter()
            dt = t1 - t0
            if best is None or dt < best[0]:
                best = (dt, int(view.pos_ptr[view.n_rows]))
            pbio.lib().lbio_rows_free(ctypes.byref(view))
        decode_s, n_cells = best
        decoded = pbio.rows_read(path, mf)
        eng = Engine(abi, device=device)
        eng.set_model(model)
        decoded.upload(eng)        # warm-up (allocations)
        eng.sync()
        t0 = time.perf_counter()
        decoded.upload(eng)
        eng.sync()
        upload_s = time.perf_counter() - t0
        eng.close()
        return {"rows": int(n_msgs), "cells": n_cells, "file_bytes": file_bytes, "host_threads": os.cpu_count(),
                "decode_ms": decode_s * 1e3, "rows_per_s": n_msgs / decode_s, "cells_per_s": n_cells / decode_s,
                "file_MB_per_s": file_bytes / decode_s / 1e6,
                "upload_and_expand_ms": upload_s * 1e3, "upload_cells_per_s": n_cells / upload_s,
                "note": "rows.pbs.gz -> CSR diff on the host (lbio_rows_read), then lb_set_rows_diff expands it "
                        "into the columnar block on the device"}
    finally:
        os.unlink(path)


def measure_loom_infer(model, rows, device=0, exe=None, extra_passes=1.0, timeout=90):
    """The process boundary: `loom_infer` -- the binary loom.runner starts -- on the workload's
    files: config.pb.gz, rows.pbs.gz, model.pb.gz in; model, groups, assignments, log out.  One chain, cat
    kernel only, `extra_passes` extra passes of the annealing schedule.  Host wall clock around the whole
    process: CUDA context, file decode, upload, every batch's sweep, dump, file encode."""
    import shutil
    from loom_b200 import infer as driver
    from loom_b200 import pbio
    exe = exe or os.path.join(REPO, "loom_b200", "host", "loom_infer")
    d = tempfile.mkdtemp(prefix="loom_b200_infer_")
    try:
        mf = pbio.ModelFile.create(model)
        mf.write(os.path.join(d, "model.pb.gz"))
        pbio.rows_write(os.path.join(d, "rows.pbs.gz"), mf, pbio.Rows.from_block(rows))
        c = pbio.config_defaults()
        c.seed, c.extra_passes, c.kind_iterations, c.hyper_run = 1, extra_passes, 0, 0
        pbio.config_write(os.path.join(d, "config.pb.gz"), c)
        # the schedule depends only on counts: replay it to know how many actions the run performs
        R = rows.n_rows
        accelerating = driver.AcceleratingSchedule(c.extra_passes, c.small_data_size, c.big_data_size)
        annealing = driver.AnnealingSchedule(accelerating.extra_passes(0))
        batching, window = driver.BatchingSchedule(0), driver.RowWindow(R)
        assigned, n_actions, batches = 0, 0, 0
        while assigned != R:
            acts, assigned, boundary = driver.next_batch(R, assigned, annealing, batching, window)
            n_actions += len(acts)
            if not boundary:
                break
            batches += 1
            annealing.set_extra_passes(accelerating.extra_passes(assigned))
        out = [os.path.join(d, n) for n in ("out.model.pb.gz", "out.groups", "out.assign.pbs.gz")]
        cmd = [exe, os.path.join(d, "config.pb.gz"), os.path.join(d, "rows.pbs.gz"), "--none",
               os.path.join(d, "model.pb.gz"), "--none", "--none", "--none", out[0], out[1], out[2], "--none", "--none"]
        env = dict(os.environ, LOOM_B200_DEVICE=str(device), LOOM_B200_TIMING="1")
        t0 = time.perf_counter()
        proc = subprocess.run(cmd, capture_output=True, timeout=timeout, env=env)
        wall = time.perf_counter() - t0
        if proc.returncode != 0:
            return {"error": proc.stderr.decode(errors="replace")[-300:]}
        n_assigned = len(pbio.assign_read(out[2])[0])
        cells = n_actions * rows.n_cells() / float(R)
        phases = [l.strip() for l in proc.stderr.decode(errors="replace").splitlines() if "LOOM_B200_TIMING" in l]
        return {"wall_s": wall, "rows": R, "row_actions": n_actions, "batches": batches, "assigned_rows_out": n_assigned,
                "phases": phases,
                "cells_per_s": cells / wall, "extra_passes": extra_passes,
                "note": "whole process: context + rows.pbs.gz decode + upload + %d batch sweeps (one chain, one CTA per "
                        "kind) + dump to model/groups/assign files" % batches}
    finally:
        shutil.rmtree(d, ignore_errors=True)


_BENCH_START = time.time()
AUX_BUDGET_S = 270.0        # the auxiliary experiments stop being started once the whole run is this old


def measure_in_subprocess(script, script_args, timeout=60, env=None):
    """Run one of benchmarks/*.py in its OWN process and return the JSON line it prints: the rows of
    SURVEY.md 8f added last (query path, tare / sparsify) are measured beside the headline without being
    able to disturb it (a fault in them cannot take the bench process or its CUDA context down)."""
    cmd = [sys.executable, os.path.join(REPO, "benchmarks", script)] + [str(a) for a in script_args]
    left = AUX_BUDGET_S - (time.time() - _BENCH_START)
    if left < 15:
        return {"skipped": "time budget of the default run (%.0f s) used up" % AUX_BUDGET_S}
    try:
        proc = subprocess.run(cmd, capture_output=True, timeout=timeout, env=dict(os.environ, **(env or {})))
    except subprocess.TimeoutExpired:
        return {"error": "timed out after %.0f s" % timeout}
    if proc.returncode != 0:
        return {"error": "exit %d: %s" % (proc.returncode, proc.stderr.decode(errors="replace")[-300:])}
    try:
        return json.loads(proc.stdout.decode().strip().splitlines()[-1])
    except Exception as exc:
        return {"error": "unparsable output: %s" % exc}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.proc = None
        self.path = None
        self.gpu = gpu_index

    def start(self):
        if os.environ.get("LB_BENCH_CLOCK_MS") == "0":
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                 "-lms", os.environ.get("LB_BENCH_CLOCK_MS", "200")], stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def pin(arr):
    """cudaHostRegister the numpy buffer so the H2D copies of the e2e leg are from pinned memory."""
    try:
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
        if arr.nbytes:
            return rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0) == 0
    except OSError:
        pass
    return False


EPHEMERAL = 32            # loom/config.py:55-61 `ephemeral_kind_count`
KIND_ITERATIONS = 32      # loom/config.py `kernels.kind.iterations`


def inference_step(engines, acts, seeds, it, hyper=False, stats=None):
    """One steady-state batch of Loom::infer_multi_pass with kind structure (src/loom.cc:256-313) for every chain:
    the cat kernel's REMOVE + ADD pass (all chains in one launch), then per chain KindKernel::try_run and
    HyperKernel::run.  Returns the number of features that changed kind."""
    from loom_b200.engine import cat_sweep_multi
    t0 = time.perf_counter()
    for c, e in enumerate(engines):      # the step's ephemeral kinds are seated on host threads while the device sweeps
        e.prepare_featureless_kinds(EPHEMERAL, seed=int(seeds[c]) + 11)
    cat_sweep_multi(engines, acts, seeds)
    if stats is not None:
        for e in engines:
            e.sync()
        t1 = time.perf_counter()
        stats["cat_s"] = stats.get("cat_s", 0.0) + t1 - t0
    # the chains' kind kernels run side by side (one host thread per chain, as loom runs one process per sample): one
    # chain's block Pitman-Yor sampler / ephemeral-kind seating on the host overlaps another chain's K4 / K5 on the device
    def kind_kernel(c):
        e = engines[c]
        ta = time.perf_counter()
        e.kind_proposer_score(fetch=False)                       # K4 + K5
        tb = time.perf_counter()
        _, n = e.kind_run(KIND_ITERATIONS, seed=int(seeds[c]) + 7)   # block Pitman-Yor + move_features
        tc = time.perf_counter()
        e.init_featureless_kinds(EPHEMERAL, seed=int(seeds[c]) + 11)
        return n, tb - ta, tc - tb, time.perf_counter() - tc

    if len(engines) > 1:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(engines), 16)) as pool:
            results = list(pool.map(kind_kernel, range(len(engines))))
    else:
        results = [kind_kernel(0)]
    moved = sum(r[0] for r in results)
    if stats is not None:
        for key, i in (("kind_score_s", 1), ("kind_run_s", 2), ("kind_ephemeral_s", 3)):
            stats[key] = stats.get(key, 0.0) + sum(r[i] for r in results) / len(results)     # mean per chain
    if stats is not None:
        for e in engines:
            e.sync()
        t2 = time.perf_counter()
        stats["kind_s"] = stats.get("kind_s", 0.0) + t2 - t1
    if hyper:
        for c, e in enumerate(engines):
            e.hyper_run(seed=int(seeds[c]) + 13)
        if stats is not None:
            for e in engines:
                e.sync()
            stats["hyper_s"] = stats.get("hyper_s", 0.0) + time.perf_counter() - t2
    return moved


def make_chain(abi, model, rows_or_owner, grids, device=0, share=False):
    from loom_b200.engine import Engine
    e = Engine(abi, device=device)
    e.set_model(model)
    if share:
        e.share_rows(rows_or_owner)
    else:
        e.set_rows(rows_or_owner)
    e.set_hyper_prior(grids)
    return e


def run_reference_arm(args, model, rows, probe_only=False):
    """Restated reference on the host cores: the same step on a bounded row sample, as many chains in parallel as
    there are cores (loom/tasks.py:192 parallel_map over samples)."""
    from loom_b200 import _abi, hyperprior, sweep_actions
    from loom_b200.engine import RowBlock
    libs = {}
    for name in ("libloom_oracle.so", "libloom_oracle_fastmath.so"):
        path = os.path.join(REPO, "oracle", name)
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
        libs[name] = _abi.Abi(path, "lo_")
    cores = os.cpu_count() or 1
    grids = hyperprior.defaults()

    def sample(R):
        R = min(rows.n_rows, R)
        W = (R + 31) // 32
        sub = RowBlock(R, rows.cat8[:, :R].copy(), rows.wide[:, :R].copy(), rows.mask[:, :W].copy())
        if R % 32:
            sub.mask[:, -1] &= np.uint32((1 << (R % 32)) - 1)
        return sub

    def measure(lib, sub, n_par, threads_per_chain, warmup, steps):
        """n_par chains side by side, each with `threads_per_chain` OpenMP threads (one per kind in the cat kernel,
        src/cat_pipeline.cc:103-125; over features in the kind / hyper kernels, src/kind_proposer.cc:339)"""
        lib.lib.lo_set_threads(threads_per_chain)
        R = sub.n_rows
        acts = sweep_actions(R, add=True, remove=True)
        engines = []
        for c in range(n_par):
            e = make_chain(lib, model, sub, grids)
            e.init_featureless_kinds(EPHEMERAL, seed=100 + c)
            engines.append(e)

        def chain(c, it):
            if it < 0:
                engines[c].cat_sweep(sweep_actions(R), seed=50 + c)          # the greedy first pass (untimed set-up)
            else:
                inference_step([engines[c]], acts, [1000 * c + it], it, hyper=args.hyper)

        def run_all(it):
            t0 = time.perf_counter()
            ths = [threading.Thread(target=chain, args=(c, it)) for c in range(n_par)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            return time.perf_counter() - t0

        run_all(-1)
        times = [run_all(it) for it in range(warmup + steps)][warmup:]
        for e in engines:
            e.close()
        ms = 1e3 * float(np.mean(times))
        return n_par * 2 * sub.n_cells() / (ms * 1e-3), ms

    # which build / threading is the fastest is decided on a small probe; the timed run uses the winner only
    probe = sample(min(args.ref_rows, 4000))
    k_threads = min(cores, 8)
    candidates = [(name, n, t) for name in libs for n, t in ((cores, 1), (max(1, cores // k_threads), k_threads))]
    probed = [(measure(libs[name], probe, n, t, 0, 1)[0], name, n, t) for name, n, t in candidates]
    _, name, n_par, tpc = max(probed)
    sub = sample(args.ref_rows)
    value, ms = measure(libs[name], sub, n_par, tpc, 0 if probe_only else args.warmup, 1 if probe_only else args.steps)
    threads = min(cores, n_par * tpc)
    note = "; ".join("%s %d x %d: %.3g" % (nm.replace("libloom_oracle", "oracle").replace(".so", ""), n, t, v) for v, nm, n, t in probed)
    return {
        "value": value, "ms_per_step": ms,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "host_cores": cores, "chains_in_parallel": n_par, "threads_per_chain": tpc, "build": name,
                         "sample": "the same step (REMOVE+ADD pass over the rows, kind proposal round) on the "
                                   "first %d of %d rows of %s, %d chain(s) side by side; float64 C restatement of the "
                                   "reference (distributions 2.0.28 closed forms); build and threading picked by a "
                                   "%d-row probe, cells/s: [%s]" % (sub.n_rows, rows.n_rows, args.workload, n_par,
                                                                    probe.n_rows, note)},
    }


def shard_bytes(model, e, n_kinds):
    """table bytes one K5 pass reads: uint32 [V_f + 1][pst] per DD feature per kind"""
    total = 0
    for k in range(n_kinds):
        pst = (e.kind_info(k).n_groups + 31) // 32 * 32
        for f in model.features:
            rows_i = {"bb": 2, "gp": 2, "nich": 1}.get(f["type"]) or len(f.get("alphas", f.get("betas", ()))) + 1
            total += 4 * rows_i * pst
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--rows", type=int, default=0, help="override the row count (parity / smoke runs only)")
    ap.add_argument("--ref-rows", type=int, default=20000, help="rows of the CPU legs' bounded sample")
    ap.add_argument("--samples", type=int, default=10, help="chains per GPU (loom's sample_count, default 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--hyper", action="store_true", help="HyperKernel::run inside the step too")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    config = {"workload": args.workload, "rows": None, "features": None, "kinds": None, "ephemeral_kinds": EPHEMERAL,
              "samples_per_gpu": args.samples,
              "step": "per chain: REMOVE+ADD of every row (cat kernel, K+32 kinds) + KindKernel::try_run (K4, K5, %d block "
                      "Pitman-Yor iterations, move_features, fresh ephemeral kinds)%s" % (
                          KIND_ITERATIONS, " + HyperKernel::run (default grids)" if args.hyper else "")}

    if args.impl == "reference":
        if rank != 0:
            return 0
        model, rows = load_workload(args.workload, args.rows)
        config.update(rows=rows.n_rows, features=model.n_features, kinds=model.n_kinds)
        res = run_reference_arm(args, model, rows)
        config["parallelism"] = "host cores only: %d chain(s) x %d thread(s) (see cpu_baseline.sample)" % (
            res["cpu_baseline"]["chains_in_parallel"], res["cpu_baseline"]["threads_per_chain"])
        config["l2_policy"] = "n/a (CPU)"
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": res["cpu_baseline"],
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    from loom_b200 import _abi, dist, hyperprior, sweep_actions
    from loom_b200.engine import cat_sweep_multi
    cuda = _abi.load_cuda()  # fails loudly if the CUDA library is missing
    if rank == 0:
        model, rows, truth = load_workload(args.workload, args.rows, with_truth=True)
    grids = hyperprior.defaults()
    S = args.samples

    # chain 0 of every rank doubles as the communicator's engine (one NCCL communicator per rank, made by the library)
    if rank != 0:     # rank 0 generated (or found) the cache; the others wait for the file instead of regenerating
        from loom_b200 import synth
        cfg = dict(synth.CONFIGS[args.workload])
        cache = os.path.join(tempfile.gettempdir(), "loom_b200_%s_%d_%d_v2.npz" % (args.workload, args.rows or cfg["rows"], cfg["seed"]))
        t0 = time.time()
        while not os.path.exists(cache) and time.time() - t0 < 600:
            time.sleep(0.5)
        model, rows, truth = load_workload(args.workload, args.rows, with_truth=True)
    R = rows.n_rows
    n_obs = rows.n_cells()
    config.update(rows=R, features=model.n_features, kinds=model.n_kinds)
    for a in (rows.cat8, rows.wide, rows.mask):
        pin(a)

    engines = []
    for c in range(S):
        e = make_chain(cuda, model, rows if c == 0 else engines[0], grids, device=local_rank, share=c > 0)
        e.init_featureless_kinds(EPHEMERAL, seed=1000 * (rank * S + c) + 1)
        engines.append(e)
    e = engines[0]
    if world > 1:
        dist.init(e, rank, world)
    K_all = e.n_kinds()
    config["parallelism"] = ("%d independent chains per GPU x %d kinds (%d with features + %d ephemeral) in ONE launch "
                             "per sweep (lb_cat_sweep_multi): one CTA per (chain, kind); all chains of a GPU read one "
                             "shared row block%s" % (S, K_all, model.n_kinds, EPHEMERAL,
                                                     "; x%d ranks, chains independent (replicas)" % world if world > 1 else ""))
    config["l2_policy"] = ("inputs larger than L2: row block %.0f MB + per-chain tables (%d chains x ~%.0f MB) against a "
                           "126 MB L2; every step rewrites all tables" % (
                               (rows.cat8.nbytes + rows.wide.nbytes + rows.mask.nbytes) / 1e6, S,
                               sum(8.0 * 64 * (len(f.get("alphas", ())) + 3) for f in model.features) / 1e6))

    def sync_all():
        for eng in engines:
            eng.sync()

    def barrier():
        sync_all()
        if world > 1:
            dist.barrier(e)

    def max_over_ranks(x):
        return float(dist.max_over_ranks(e, [x])[0]) if world > 1 else x

    def seeds_for(it):
        return np.array([1000 * (rank * S + c) + 17 * (it + 1) + 3 for c in range(S)], np.uint64)

    pass_acts = sweep_actions(R, add=True, remove=True)
    # ---- set-up (untimed): the greedy first pass of every chain (Loom::infer_single_pass) ----
    t_setup = time.perf_counter()
    cat_sweep_multi(engines, sweep_actions(R), seeds_for(-1))
    sync_all()
    log("rank %d: first pass of %d chains over %d rows in %.1f s" % (rank, S, R, time.perf_counter() - t_setup))

    # ---- device-resident value ----
    for it in range(args.warmup):
        inference_step(engines, pass_acts, seeds_for(it), it, hyper=args.hyper)
    barrier()
    for eng in engines:
        eng.launch_count(reset=True)
        eng.kernel_ms(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    phases, moved = {}, 0
    t0 = time.perf_counter()
    e.timer_start()                            # CUDA events on the stream the sweeps are launched on
    for it in range(args.steps):
        moved += inference_step(engines, pass_acts, seeds_for(args.warmup + it), it, hyper=args.hyper, stats=phases)
    sync_all()                                 # every chain's stream has drained before the stop event is recorded
    dev_ms = e.timer_stop()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(eng.launch_count() for eng in engines)
    kms = sum(eng.kernel_ms() for eng in engines)
    total_ms = max_over_ranks(max(dev_ms, wall_ms))   # device events / host view, max over ranks
    ms_per_step = total_ms / args.steps
    cells_per_step = 2 * n_obs * S
    value = world * cells_per_step / (ms_per_step * 1e-3)
    cat_ms = max_over_ranks(phases.get("cat_s", 0.0) * 1e3) / args.steps
    kind_ms = max_over_ranks(phases.get("kind_s", 0.0) * 1e3) / args.steps
    hyper_ms = max_over_ranks(phases.get("hyper_s", 0.0) * 1e3) / args.steps
    F = model.n_features

    # ---- e2e: rows streamed in from pinned host memory, assignments read back, every step ----
    def e2e_step(it):
        e.reload_rows(rows)                   # H2D of the whole row block (the other chains borrow it)
        inference_step(engines, pass_acts, seeds_for(100 + it), it, hyper=args.hyper)
        n = 0
        for eng in engines:
            for k in range(eng.n_kinds() - EPHEMERAL):
                eng.get_assignments(k)        # D2H: the assignment column of every feature-bearing kind
                n += 1
        return n

    barrier()
    t0 = time.perf_counter()
    cols = 0
    e2e_steps = max(1, min(args.steps, 2))
    for it in range(e2e_steps):
        cols += e2e_step(it)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = world * cells_per_step / (e2e_ms * 1e-3)
    h2d = int(rows.cat8.nbytes + rows.wide.nbytes + rows.mask.nbytes + 4 * len(pass_acts))
    d2h = int(4 * R * cols / e2e_steps)

    # ---- roofline of the dominant kernel ----
    peaks = {}
    ppath = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # one "launch" of the dominant kernel = one sweep call: the four class kernels (16-warp, 8-warp, one-warp, featureless)
    # run side by side on four streams and the 16-warp class (k_cat_sweep_spec<16>) is the longest; sweep_ms is the device
    # time from the first kernel's start to the last one's end, relaunch rounds after table growth included
    n_sweep = args.steps
    sweep_ms = float(kms[1])
    bytes_per_action = algorithmic_bytes_per_action(model, rows, K_all)
    alg_bytes_per_launch = bytes_per_action * len(pass_acts) * S
    avg_launch_ms = sweep_ms / max(n_sweep, 1)
    achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms else 0.0
    traffic = None
    tpath = os.path.join(REPO, "profiles", "r2_k_cat_sweep_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("workload") == args.workload and t.get("chains") == S and t.get("rows") == R:
            traffic = t.get("dram_bytes_per_launch")           # ncu dram__bytes_{read,write}.sum of one sweep at this size
    roofline = {"bound": "hbm", "kernel": "k_cat_sweep_spec", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": avg_launch_ms, "launches": n_sweep, "class_kernels_launched": int(launches[1]),
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "kernel_share_of_step": sweep_ms / (dev_ms or 1.0),
                "chains_per_launch": S, "algorithmic_bytes_per_row_action": bytes_per_action,
                "note": "exact sequential Gibbs: a dependent chain per (chain, kind) CTA, bound by the latency of one "
                        "sample -> update -> rescore-one-column round per row action, not by bandwidth"}

    # ---- the second half of the metric: kind-kernel feature-scores/s ----
    kind = {"metric": "kind_kernel_feature_scores_per_sec", "unit": "feature-scores/s",
            "features": F, "kinds_scored": K_all,
            "in_step": {"value": world * S * F * K_all / (kind_ms * 1e-3) if kind_ms else None, "ms_per_step": kind_ms,
                        "note": "KindKernel::try_run of every local chain inside the timed step: K4 + K5 + %d block "
                                "Pitman-Yor iterations (host) + move_features + new ephemeral kinds" % KIND_ITERATIONS}}
    aux = {"phases_ms_per_step": {"cat_kernel": cat_ms, "kind_kernel": kind_ms, "hyper_kernel": hyper_ms,
                                  "kind_kernel_parts_mean_per_chain_rank0": {k_: 1e3 * v_ / args.steps for k_, v_ in phases.items() if k_.startswith("kind_") and k_ != "kind_s"}},
           "features_moved_per_step": moved / float(args.steps)}
    try:
        # ONE chain, identical on every rank (the generating partition), its proposal round sharded over the ranks
        sh = make_chain(cuda, model, e, grids, device=local_rank, share=True)
        install_truth(sh, truth)
        sh.init_featureless_kinds(EPHEMERAL, seed=5)
        if world > 1:
            dist.init(sh, rank, world, tag="sh%s" % os.environ.get("MASTER_PORT", "0"))
        Ks = sh.n_kinds()
        dist.sharded_kind_proposer_score(sh, fetch=False)          # warm-up (allocations, NCCL channels)
        reps, acc = 5, {}
        for _ in range(reps):
            if world > 1:
                dist.barrier(sh)
            t0 = time.perf_counter()
            _, st = dist.sharded_kind_proposer_score(sh, fetch=True)
            st["round_wall_ms"] = (time.perf_counter() - t0) * 1e3
            for k_, v_ in st.items():
                acc.setdefault(k_, []).append(v_)
        names = ("accumulate_ms", "allreduce_ms", "score_ms", "allgather_ms", "round_wall_ms")
        mx = dist.max_over_ranks(sh, [float(np.median(acc[n])) for n in names]) if world > 1 else [float(np.median(acc[n])) for n in names]
        st = dict(zip(names, [float(x) for x in mx]))
        dev_round = st["accumulate_ms"] + st["allreduce_ms"] + st["score_ms"] + st["allgather_ms"]
        tbl = shard_bytes(model, sh, Ks)
        ar_bytes, ag_bytes = float(acc["allreduce_bytes"][0]), float(acc["allgather_bytes"][0])
        kind["sharded"] = {
            "scaling": "strong", "ranks": world, "value": F * Ks / (dev_round * 1e-3), "round_ms": dev_round,
            "round_wall_ms": st["round_wall_ms"], "k4_accumulate_ms": st["accumulate_ms"], "k5_score_ms": st["score_ms"],
            "allreduce_ms": st["allreduce_ms"], "allreduce_bytes": ar_bytes,
            "allreduce_busbw_GBs": (2.0 * (world - 1) / world * ar_bytes / (st["allreduce_ms"] * 1e-3) / 1e9) if world > 1 and st["allreduce_ms"] else None,
            "allgather_ms": st["allgather_ms"], "allgather_bytes": ag_bytes,
            "k4_proposer_cells_per_s": n_obs * Ks / (st["accumulate_ms"] * 1e-3) * 1.0,
            "k5_alone_feature_scores_per_s": F * Ks / (st["score_ms"] * 1e-3),
            "note": "one chain (the generating partition, identical on every rank), one proposal round: rows sharded for "
                    "K4, one ncclAllReduce per slab, features sharded for K5, one ncclAllGather; device ms, max over ranks"}
        kind["roofline_k5"] = {"bound": "hbm", "kernel": "k_prop_score", "achieved": tbl / world / (st["score_ms"] * 1e-3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": tbl / world / (st["score_ms"] * 1e-3) / 1e9 / peak,
                               "table_bytes": tbl, "note": "count tables read once per round (SURVEY.md 8d: stat bytes x G "
                               "per feature-score); score_ms also holds the two small per-round table kernels"}
        k4_bytes = (bytes_per_action - 8.0 * K_all - 8.0) * R * 1.0 / world * Ks + 1.0 * R / world * Ks
        kind["roofline_k4"] = {"bound": "hbm", "kernel": "k_accum_feature", "achieved": k4_bytes / (st["accumulate_ms"] * 1e-3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": k4_bytes / (st["accumulate_ms"] * 1e-3) / 1e9 / peak,
                               "note": "per row x kind: the row's cells + mask + 1 B packed group id (SURVEY.md 8d); the "
                                       "kernel is bound by shared-memory atomics, one per observed cell x kind"}
        # task-sharded hyper kernel on the same chain
        dist.sharded_hyper_run(sh, seed=3)
        sh.sync()
        t0 = time.perf_counter()
        ar_ms = dist.sharded_hyper_run(sh, seed=4)
        sh.sync()
        hy = max_over_ranks((time.perf_counter() - t0) * 1e3)
        aux["hyper_kernel_sharded"] = {"ranks": world, "ms": hy, "allreduce_ms": ar_ms, "features_per_s": F / (hy * 1e-3),
                                       "note": "HyperKernel::run, task t on rank t %% world, one all-reduce of the choices"}
        if world > 1:
            # kind-sharded sweep of that chain + publication of the kinds (broadcast group)
            owner = np.arange(Ks) % world
            n_act = min(R, 50000)
            sh.cat_sweep(sweep_actions(n_act, add=True, remove=True), seed=9, kind_mask=(owner == rank).astype(np.uint8))
            nbytes, ms = dist.sync_kinds(sh, owner)
            ms = max_over_ranks(ms)
            aux["kind_sharded_sweep_publish"] = {"ranks": world, "bytes": nbytes, "ms": ms, "GBs": nbytes / (ms * 1e-3) / 1e9 if ms else None,
                                                 "note": "after a kind-sharded sweep every kind's tables, id maps and assignment "
                                                         "column are broadcast from its owner (one NCCL group)"}
        sh.close()
    except Exception as exc:  # the auxiliary numbers never mask the headline
        kind["sharded"] = {"error": str(exc)}

    # ---- auxiliary: the bandwidth-shaped batch kernels on the same workload, rank 0 ----
    if rank == 0 and not args.no_aux:
        try:
            # K1 batch score_value, scores materialised, the largest kind of chain 0
            f2k_now = e.get_kinds()[1]
            big = int(np.argmax(np.bincount(f2k_now, minlength=e.n_kinds())))
            kf = np.nonzero(f2k_now == big)[0]
            G = e.kind_info(big).n_groups
            for _ in range(2):
                e.score_rows(big, fetch=False)
            e.sync()
            e.timer_start()
            reps = 5
            for _ in range(reps):
                e.score_rows(big, fetch=False)
            ms = e.timer_stop() / reps
            width = np.array([1 if model.features[f]["type"] in ("bb", "dd") else 4 for f in kf], np.float64)
            cells_k = rows.cells_per_feature()[kf].astype(np.float64)
            in_bytes = float((cells_k * width).sum()) + R * len(kf) / 8.0
            out_bytes = 4.0 * G * R
            gbs = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9
            aux["k1_score_rows"] = {"kind": big, "features": len(kf), "groups": G, "ms": ms,
                                    "cells_per_s": float(cells_k.sum()) / (ms * 1e-3), "achieved": gbs,
                                    "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bound": "hbm"}
            # K2 bulk accumulate of all kinds with the current assignments (remove then add back)
            e.accumulate_rows(sign=-1)
            e.accumulate_rows(sign=+1)
            e.sync()
            e.timer_start()
            e.accumulate_rows(sign=-1)
            e.accumulate_rows(sign=+1)
            ms2 = e.timer_stop() / 2
            in_all = bytes_per_action * R
            aux["k2_accumulate_rows"] = {"ms": ms2, "cells_per_s": n_obs / (ms2 * 1e-3),
                                         "achieved": in_all / (ms2 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                         "frac": in_all / (ms2 * 1e-3) / 1e9 / peak, "bound": "hbm"}
        except Exception as exc:
            aux["error"] = str(exc)

    if rank == 0 and world == 1 and not args.no_aux:
        # the other single-GPU configuration's shape (C4, taxi-like): score_value of the decoded rows against score_diff from the
        # tare-compressed form, and the tensor-core form of K1 on the kinds it is meant for -- each in a process of its own
        aux["c4_k1_dense_vs_sparse_score_diff"] = measure_in_subprocess("k1d_bench.py", [200000], timeout=150)
        aux["k1_tcgen05_vs_gather_bb256"] = measure_in_subprocess("k1_bench.py", ["bb", 256, 1000000], timeout=120)
        aux["k1_tcgen05_vs_gather_dd16"] = measure_in_subprocess("k1_bench.py", ["dd16", 128, 1000000, 2], timeout=120)

    # ---- rates at other chain counts (rank 0, N = 1): one chain alone; 100 chains on a 100 k-row block ----
    rates = {"cells_per_step_per_chain": 2 * n_obs, "observed_cells": n_obs, "chains": S,
             "row_actions_per_s": world * S * len(pass_acts) / (ms_per_step * 1e-3),
             "cat_kernel_only_cells_per_s": world * cells_per_step / (cat_ms * 1e-3) if cat_ms else None}
    if rank == 0 and not args.no_aux:
        try:
            e.sync()
            e.timer_start()
            e.cat_sweep(pass_acts[:min(len(pass_acts), 400000)], seed=77)
            ms1 = e.timer_stop()
            n1 = min(len(pass_acts), 400000)
            rates["single_chain_cells_per_s"] = n_obs / R * n1 / (ms1 * 1e-3)
            rates["single_chain_us_per_row_action"] = 1e3 * ms1 / n1
        except Exception as exc:
            rates["single_chain_error"] = str(exc)
        if world == 1:
            for n_ch, n_rows in ((100, 100000), (1008, 20000), (30, 200000)):
                r = measure_in_subprocess("sweep_bench.py", [args.workload, n_ch, n_rows, EPHEMERAL], timeout=150,
                                          env={"LB_BENCH_CLOCK_MS": "0"})
                rates["chains_%d" % n_ch] = ({"cells_per_s": r.get("phases", {}).get("remove_add_pairs", {}).get("cells_per_s"),
                                              "rows": n_rows, "note": "REMOVE+ADD pass, %d chains in one launch" % n_ch}
                                             if "phases" in r else r)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches[0]),
            "roofline": roofline, "kind_kernel": kind, "roofline_aux": aux, "clocks": clocks, "rates": rates}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = run_reference_arm(args, model, rows, probe_only=True)["cpu_baseline"]
        except Exception as exc:
            line["cpu_baseline"] = {"error": str(exc)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier(e)
    return 0


if __name__ == "__main__":
    sys.exit(main())
