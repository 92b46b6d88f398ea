#!/usr/bin/env python
"""bench.py -- cat-kernel cells/sec on BASELINE.json configs[1] (C2).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload "C2"): synthetic 100k rows x 100 mixed features
(25 BB / 25 DD-256 / 25 GP / 25 NICH), density 0.5, fixed structure (7 kinds),
cat kernel only.  One STEP is one self-contained pass of the cat kernel over the
whole row block, in loom's action semantics (src/cat_kernel.hpp:177-299):
    ADD every row (greedy pass, Loom::infer_single_pass)          R adds
    REMOVE+ADD every row (one extra-pass refresh, loom.cc:502-505) R removes + R adds
    REMOVE every row (drain; returns the mixtures to empty)        R removes
so every step starts from the same empty state and does identical work.  A
"cell" is one observed (row, feature) value consumed by one cat-kernel action
(SURVEY.md 8d): a step consumes 4 x (#observed cells).

value   = cells/s with the row block already resident in HBM (CUDA events).
e2e     = the same step through the C ABI from HOST buffers: the row block is
          re-uploaded from pinned host memory and the assignments of every kind
          are read back inside the timed region.
roofline= the dominant kernel (k_cat_sweep): algorithmic bytes per launch
          (in + 4K read + 4K written per row-action) / its CUDA-event
          time, against the measured HBM copy bandwidth.  The exact sequential
          Gibbs chain is latency-bound (one CTA per kind); roofline_aux carries the
          batch kernels (score_rows K1, accumulate_rows K2) that ARE HBM-bound.
samples : the exact Gibbs chain of one kind is sequential, so one chain keeps only
          K (=7) of the 148 SMs busy.  loom itself runs `sample_count` independent
          chains per dataset in parallel processes (loom/tasks.py:173-194, default 10)
          over ONE rows file; the GPU sweeps many chains in ONE launch per action list
          (lb_cat_sweep_multi, rows shared with lb_share_rows): one CTA per (chain, kind),
          one warp wide when there are enough chains to fill every SM 16 times over
          (default 1008 chains = three waves), 8 warps wide for a single chain.
          `value` is the aggregate over those chains ("config.samples"); the
          single-chain rate is in "rates".  The reference arm gets the same treatment:
          as many chains in parallel as the host cores allow.
N > 1  : chains do not communicate, so every rank runs its own set of chains on its
          GPU -- "replicas only", "scaling": "weak" (DESIGN.md).
roofline_aux (N = 1, rank 0) also carries, each measured in a process of its own (scratch/*_bench.py) so that it
          cannot disturb the headline and only started while the run is younger than AUX_BUDGET_S: the query path
          (`query`, `query_fused`), the tare / sparsify passes (`differ`), the per-feature-type score error against the
          float64 oracle (`score_error`), and A/Bs of the sweep kernel for the next round
          (`k3_variant_score_from_counts`: product library, -DLB_SCORE_FROM_COUNTS build, compact tables).
--impl reference times the restated reference (oracle/, float64 C, OpenMP one
thread per kind like src/cat_pipeline.cc:103-125) on the host cores, on a bounded
row sample of the same workload.
"""
import argparse
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


def load_workload(name, rows):
    from loom_b200 import synth
    from loom_b200.engine import Model, RowBlock
    cfg = dict(synth.CONFIGS[name])
    if rows:
        cfg["rows"] = rows
    cache = os.path.join(tempfile.gettempdir(), "loom_b200_%s_%d_%d.npz" % (name, cfg["rows"], cfg["seed"]))
    feats = synth.make_features(cfg["types"])
    if os.path.exists(cache):
        z = np.load(cache)
        model = Model(feats, z["f2k"], kind_py=[synth.CLUSTERING] * (int(z["f2k"].max()) + 1),
                      topology=synth.CLUSTERING)
        return model, RowBlock(cfg["rows"], z["cat8"], z["wide"], z["mask"])
    model, rows_, _ = synth.generate(**cfg)
    try:
        np.savez(cache, f2k=model.f2k, cat8=rows_.cat8, wide=rows_.wide, mask=rows_.mask)
    except OSError:
        pass
    return model, rows_


def step_actions(R):
    from loom_b200 import sweep_actions
    first = np.concatenate([sweep_actions(R), sweep_actions(R, add=True, remove=True)])
    drain = sweep_actions(R, add=False)
    return first, drain


def algorithmic_bytes_per_action(model, rows):
    """per row-action, input cells (BB 1 B, DD 1 B, DPD/GP/NICH 4 B
    per observed cell) + 1 bit/feature mask + 8 B row id, + 4 B/kind read and
    4 B/kind written (assignment ids)."""
    obs = rows.observed()
    width = np.array([1 if f["type"] in ("bb", "dd") else 4 for f in model.features])
    cell_bytes = float((obs * width[:, None]).sum()) / rows.n_rows
    return cell_bytes + model.n_features / 8.0 + 8.0 + 8.0 * model.n_kinds


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
        env = dict(os.environ, LOOM_B200_DEVICE=str(device))
        t0 = time.perf_counter()
        proc = subprocess.run(cmd, capture_output=True, timeout=timeout, env=env)
        wall = time.perf_counter() - t0
        if proc.returncode != 0:
            return {"error": proc.stderr.decode(errors="replace")[-300:]}
        n_assigned = len(pbio.assign_read(out[2])[0])
        cells = n_actions * rows.n_cells() / float(R)
        return {"wall_s": wall, "rows": R, "row_actions": n_actions, "batches": batches, "assigned_rows_out": n_assigned,
                "cells_per_s": cells / wall, "extra_passes": extra_passes,
                "note": "whole process: context + rows.pbs.gz decode + upload + %d batch sweeps (one chain, one CTA per "
                        "kind) + dump to model/groups/assign files" % batches}
    finally:
        shutil.rmtree(d, ignore_errors=True)


_BENCH_START = time.time()
AUX_BUDGET_S = 270.0        # the auxiliary experiments stop being started once the whole run is this old


def measure_in_subprocess(script, script_args, timeout=60, env=None):
    """Run one of scratch/*_bench.py in its OWN process and return the JSON line it prints: the rows of
    SURVEY.md 8f added last (query path, tare / sparsify) are measured beside the headline without being
    able to disturb it (a fault in them cannot take the bench process or its CUDA context down)."""
    cmd = [sys.executable, os.path.join(REPO, "scratch", script)] + [str(a) for a in script_args]
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


def run_reference_arm(args, model, rows, n_cells_full):
    """Restated reference on host cores: same step, bounded row sample."""
    from loom_b200 import _abi
    from loom_b200.engine import Engine, RowBlock
    import __graft_entry__
    path = os.path.join(REPO, "oracle", "libloom_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
    oracle = _abi.Abi(path, "lo_")
    R = min(rows.n_rows, args.ref_rows)
    W = (R + 31) // 32
    sub = RowBlock(R, rows.cat8[:, :R], rows.wide[:, :R], rows.mask[:, :W])
    if R % 32:
        sub.mask[:, -1] &= np.uint32((1 << (R % 32)) - 1)
    cells = 4 * sub.n_cells()
    first, drain = step_actions(R)
    cores = os.cpu_count() or 1

    def measure(n_par, threads_per_chain):
        """n_par chains in parallel (loom/tasks.py:192 parallel_map over samples), each with
        `threads_per_chain` OpenMP threads over its kinds (src/cat_pipeline.cc:103-125)."""
        oracle.lib.lo_set_threads(threads_per_chain)
        engines = []
        for c in range(n_par):
            e = Engine(oracle)
            e.set_model(model)
            e.set_rows(sub)
            engines.append(e)

        def chain(c, it):
            engines[c].cat_sweep(first, seed=100 * c + it)
            engines[c].cat_sweep(drain, seed=100 * c + it)

        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            ths = [threading.Thread(target=chain, args=(c, it)) for c in range(n_par)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        for e in engines:
            e.close()
        ms = 1e3 * float(np.mean(times))
        return n_par * cells / (ms * 1e-3), ms

    # Two ways to use every host core; the better one is the baseline.  (a) the reference's own
    # threading: one thread per kind inside a chain, as many chains as fit.  (b) one chain per
    # core with its kinds in sequence: no thread idles behind the largest kind.
    modes = [(max(1, cores // model.n_kinds), model.n_kinds), (cores, 1)]
    results = [measure(n, t) + (n, t) for n, t in modes]
    value, ms, n_par, tpc = max(results)
    threads = min(cores, n_par * tpc)
    both = "; ".join("%d chain(s) x %d thread(s): %.3g cells/s" % (n, t, v) for v, _, n, t in results)
    return {
        "value": value, "ms_per_step": ms,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "host_cores": cores, "chains_in_parallel": n_par, "threads_per_chain": tpc,
                         "sample": "first %d of %d rows of C2, same 4-pass step; float64 C oracle (restated "
                                   "reference, distributions 2.0.28 math in exact libm); best of [%s]" % (
                                       R, rows.n_rows, both)},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--rows", type=int, default=0, help="override the row count (parity runs only)")
    ap.add_argument("--ref-rows", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=0,
                    help="independent chains swept per GPU in one launch (0 = two waves of 16 one-warp CTAs per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    config = {"workload": args.workload, "rows": None, "features": None, "kinds": None,
              "step": "ADD all + REMOVE/ADD all + REMOVE all (4 cat-kernel actions per row)",
              "l2_policy": "tables and row block (26 MB) are re-streamed every pass; every step rewrites "
                           "all per-kind tables; inputs are not larger than L2 -- the kernel is latency-bound, "
                           "not bandwidth-bound (see DESIGN.md)",
              "parallelism": "replicas x%d" % world if world > 1 else "single chain, one CTA per kind"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        model, rows = load_workload(args.workload, args.rows)
        config.update(rows=rows.n_rows, features=model.n_features, kinds=model.n_kinds)
        res = run_reference_arm(args, model, rows, 4 * rows.n_cells())
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

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    from loom_b200 import _abi
    from loom_b200.engine import Engine
    cuda = _abi.load_cuda()  # fails loudly if the CUDA library is missing
    model, rows = load_workload(args.workload, args.rows)
    R = rows.n_rows
    config.update(rows=R, features=model.n_features, kinds=model.n_kinds)
    n_obs = rows.n_cells()
    cells_per_step = 4 * n_obs
    first, drain = step_actions(R)

    from loom_b200.engine import cat_sweep_multi
    # Three waves of 16 one-warp CTAs per SM: 3 * 16 * 148 / K chains, in whole blocks of 8 chains
    # (the expensive kinds are dispatched first and the cheap ones fill the tail).
    n_chains = args.samples or max(1, (3 * 16 * 148 // model.n_kinds) // 8 * 8)
    config["samples"] = n_chains
    config["parallelism"] = ("%d independent chains x %d kinds per GPU in ONE launch per sweep "
                             "(lb_cat_sweep_multi): one CTA per (chain, kind), 1-8 warps wide by how many "
                             "chains there are; all chains read one shared row block%s" % (
                                 n_chains, model.n_kinds, "; replicas x%d" % world if world > 1 else ""))
    config["l2_policy"] = ("per-chain tables (%d chains x ~17 MB) exceed the 126 MB L2 many times over and every "
                           "step rewrites all of them; the shared 26 MB row block is re-streamed every pass" % n_chains)
    engines = []
    for c in range(n_chains):
        eng = Engine(cuda, device=local_rank)
        eng.set_model(model)
        if c == 0:
            eng.set_rows(rows)
        else:
            eng.share_rows(engines[0])       # loom's samples read one rows file (loom/tasks.py:173-194)
        engines.append(eng)
    e = engines[0]
    for a in (rows.cat8, rows.wide, rows.mask):
        pin(a)

    def barrier():
        for eng in engines:
            eng.sync()
        if dist:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def chain_seeds(base, it):
        return np.arange(n_chains, dtype=np.uint64) * 1000 + np.uint64(base * rank + it)

    def device_step(it):
        seeds = chain_seeds(100000, it)
        cat_sweep_multi(engines, first, seeds)
        cat_sweep_multi(engines, drain, seeds)

    # ---- single chain, device resident (rates.single_chain): 8-warp CTAs, lowest latency ----
    for it in range(2):
        e.cat_sweep(first, seed=it)
        e.cat_sweep(drain, seed=it)
    e.sync()
    e.timer_start()
    e.cat_sweep(first, seed=2)
    e.cat_sweep(drain, seed=2)
    single_ms = e.timer_stop()

    # ---- device-resident value: all chains, one launch per sweep ----
    for it in range(args.warmup):
        device_step(it)
    barrier()
    for eng in engines:
        eng.launch_count(reset=True)
        eng.kernel_ms(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    e.timer_start()                            # CUDA events on the stream the sweeps are launched on
    for it in range(args.steps):
        device_step(args.warmup + it)
    dev_ms = e.timer_stop()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(eng.launch_count() for eng in engines)
    kms = sum(eng.kernel_ms() for eng in engines)
    total_ms = max_over_ranks(max(dev_ms, wall_ms))   # device events / host view, max over ranks
    ms_per_step = total_ms / args.steps
    value = world * n_chains * cells_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, assignments out, inside the timed region ----
    def e2e_step(it):
        seeds = chain_seeds(500000, it)
        e.set_rows(rows)                      # H2D of the whole row block (pinned), once: chains share it
        for eng in engines[1:]:
            eng.share_rows(e)
        cat_sweep_multi(engines, first, seeds)
        for eng in engines:
            for k in range(model.n_kinds):
                eng.get_assignments(k)        # D2H of every chain's assignment columns
        cat_sweep_multi(engines, drain, seeds)

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        e2e_step(1 + it)
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(e2e_wall) / args.steps
    e2e_value = world * n_chains * cells_per_step / (e2e_ms * 1e-3)
    h2d = int(rows.cat8.nbytes + rows.wide.nbytes + rows.mask.nbytes + 4 * (len(first) + len(drain)))
    d2h = n_chains * int(4 * R * model.n_kinds)

    # ---- roofline of the dominant kernel ----
    peaks = {}
    ppath = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    n_sweep = int(launches[1])
    sweep_ms = float(kms[1])
    bytes_per_action = algorithmic_bytes_per_action(model, rows)
    actions_per_step = len(first) + len(drain)
    alg_bytes_per_launch = bytes_per_action * actions_per_step * args.steps * n_chains / max(n_sweep, 1)
    avg_launch_ms = sweep_ms / max(n_sweep, 1)
    achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms else 0.0
    traffic = None
    tpath = os.path.join(REPO, "profiles", "r1_k_cat_sweep_multi_traffic.json")
    if os.path.exists(tpath) and args.workload == "C2" and not args.rows:
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch_per_chain")   # ncu dram__bytes_{read,write}.sum
        traffic = traffic * n_chains if traffic else None
    roofline = {"bound": "hbm", "kernel": "k_cat_sweep", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": avg_launch_ms, "launches": n_sweep,
                "kernel_share_of_step": sweep_ms / total_ms if total_ms else None,
                "chains_per_launch": n_chains,
                "note": "exact sequential Gibbs: a dependent chain per (chain, kind) CTA, bound by the latency of its "
                        "table round trips, not by bandwidth; one launch sweeps all %d chains" % n_chains}

    # ---- auxiliary: the HBM-bound batch kernels on the same workload ----
    aux = {}
    try:
        e.set_rows(rows)
        e.cat_sweep(first[:R], seed=9)       # populate groups (ADD pass)
        # K1 batch score_value, scores materialised, largest kind
        big = int(np.argmax([len(model.kind_features(k)) for k in range(model.n_kinds)]))
        G = e.kind_info(big).n_groups
        for _ in range(3):
            e.score_rows(big, fetch=False)
        e.sync()
        e.timer_start()
        reps = 10
        for _ in range(reps):
            e.score_rows(big, fetch=False)
        ms = e.timer_stop() / reps
        kf = model.kind_features(big)
        obs = rows.observed()[kf]
        width = np.array([1 if model.features[f]["type"] in ("bb", "dd") else 4 for f in kf])
        in_bytes = float((obs * width[:, None]).sum()) + R * len(kf) / 8.0
        out_bytes = 4.0 * G * R
        gbs = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9
        aux["k1_score_rows"] = {"kind": big, "features": len(kf), "groups": G, "ms": ms,
                                "cells_per_s": float(obs.sum()) / (ms * 1e-3), "achieved": gbs,
                                "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bound": "hbm"}
        # K2 bulk accumulate of all kinds with the current assignments (remove then add back)
        e.accumulate_rows(sign=-1)            # warm-up (first-use allocations)
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
        # kind kernel on the same rows: 32 ephemeral kinds (loom/config.py:55-61), bulk proposer
        # accumulate (K4) + feature x kind marginal-likelihood table (K5); one feature-score =
        # one (feature, kind) entry of KindProposer's score phase (src/kind_proposer.cc:339-346)
        e.init_featureless_kinds(32, seed=11)
        Kk = e.n_kinds()
        e.kind_proposer_score(fetch=False)      # warm-up (allocations)
        e.sync()
        reps = 3
        e.timer_start()
        for _ in range(reps):
            e.kind_proposer_score(fetch=False)
        msk = e.timer_stop() / reps
        aux["kind_kernel"] = {"kinds": Kk, "features": model.n_features, "ms": msk,
                              "feature_scores_per_s": model.n_features * Kk / (msk * 1e-3),
                              "proposer_cells_per_s": n_obs * Kk / (msk * 1e-3),
                              "note": "K4 accumulate of every row into every kind + K5 scoring, per proposal round"}
        e.init_featureless_kinds(0, seed=12)
        # hyper kernel (HyperKernel::run, src/hyper_kernel.cc:195-235) with loom's default grids
        # (loom/hyperprior.py:33-71): one grid-Gibbs pass over every feature's hyper-parameters and
        # every kind's Pitman-Yor parameters
        from loom_b200 import hyperprior
        e.set_hyper_prior(hyperprior.defaults())
        e.hyper_run(seed=13)
        e.sync()
        e.timer_start()
        e.hyper_run(seed=14)
        msh = e.timer_stop()
        aux["hyper_kernel"] = {"ms": msh, "features": model.n_features,
                               "features_per_s": model.n_features / (msh * 1e-3),
                               "note": "default grids: BB 12x12, DD 12 per coordinate (256 coordinates), GP 100x100, "
                                       "NICH 201/30/100/30, Pitman-Yor 137 points"}
        e.cat_sweep(drain, seed=9)
    except Exception as exc:  # the auxiliary numbers never mask the headline
        aux["error"] = str(exc)

    if rank == 0:
        try:
            aux["ingest"] = measure_ingest(cuda, model, rows, device=local_rank)
        except Exception as exc:  # never masks the headline
            aux["ingest"] = {"error": str(exc)}
        try:
            aux["loom_infer_process"] = measure_loom_infer(model, rows, device=local_rank)
        except Exception as exc:
            aux["loom_infer_process"] = {"error": str(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # query path (8f n3) and tare / sparsify (8f n4) on the same C2 mix, each in its own process
        aux["query"] = measure_in_subprocess("query_bench.py", [R, 1000000])
        d = aux["differ"] = measure_in_subprocess("differ_bench.py", [R])
        for part in ("tare", "sparsify"):
            if isinstance(d.get(part), dict) and "achieved_gb_s" in d[part]:
                d[part].update({"achieved": d[part]["achieved_gb_s"], "peak": peak, "unit": "GB/s",
                                "frac": d[part]["achieved_gb_s"] / peak, "bound": "hbm"})
        # the same with K1 reducing scores to log-sum-exp in registers (scores not materialised)
        aux["query_fused"] = measure_in_subprocess("query_bench.py", [R, 1000], env={"LB_QUERY_FUSED": "1"})
        aux["score_error"] = measure_in_subprocess("score_error.py", [8000])     # parity check 2, per feature type
        # experiment for the next round: the sweep scoring BB / DD / DPD cells from the count rows
        # instead of the cached scorer rows -- the product library and the variant on the same reduced workload
        variant = os.path.join(REPO, "loom_b200", "csrc", "libloomb200_counts.so")
        ab = {"default": measure_in_subprocess("sweep_variant_bench.py", [168, 50000], timeout=120)}
        ab["from_counts"] = (measure_in_subprocess("sweep_variant_bench.py", [168, 50000], timeout=120, env={"LOOM_B200_LIB": variant})
                             if os.path.exists(variant) else {"error": "variant library not built (make -C loom_b200/csrc variants)"})
        # and the product library with compact tables: column capacity starting at 64 instead of 256 (a quarter of the
        # table footprint and of the per-chain shared memory; kinds that outgrow it re-stride on demand)
        ab["default_gcap64"] = measure_in_subprocess("sweep_variant_bench.py", [168, 50000], timeout=120, env={"LB_INITIAL_GCAP": "64"})
        # and with a row order per chain (lb_cat_sweep_streams: loom's per-sample shuffles) instead of the shared one
        ab["default_streams"] = measure_in_subprocess("sweep_variant_bench.py", [168, 50000, "streams"], timeout=120)
        aux["k3_variant_score_from_counts"] = ab

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches[0]),
            "roofline": roofline, "roofline_aux": aux, "clocks": clocks,
            "rates": {"cells_per_step_per_chain": cells_per_step, "observed_cells": n_obs,
                      "row_actions_per_s": world * n_chains * actions_per_step / (ms_per_step * 1e-3),
                      "single_chain_cells_per_s": cells_per_step / (single_ms * 1e-3),
                      "single_chain_ms_per_step": single_ms}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_args = argparse.Namespace(**vars(args))
        ref_args.steps, ref_args.warmup = 2, 1
        try:
            line["cpu_baseline"] = run_reference_arm(ref_args, model, rows, cells_per_step)["cpu_baseline"]
        except Exception as exc:
            line["cpu_baseline"] = {"error": str(exc)}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
