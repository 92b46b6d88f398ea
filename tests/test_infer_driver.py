"""Schedules and the whole-run driver (loom_b200/infer.py; src/schedules.hpp, src/loom.cc:167-450),
on the CPU oracle; the same driver on the GPU is in the gpu-marked test at the bottom."""
import numpy as np
import pytest

from loom_b200 import Engine, synth, hyperprior
from loom_b200 import infer


def test_annealing_rates():
    # REMOVE : ADD = N : (1 + N) (src/schedules.hpp:41-60); the first actions are ADDs
    for n in (0.0, 1.0, 3.5, 500.0):
        s = infer.AnnealingSchedule(n)
        acts = [s.next_action_is_add() for _ in range(20000)]
        adds, removes = sum(acts), len(acts) - sum(acts)
        assert acts[0] and acts[1]
        if n == 0:
            assert removes == 0
        else:
            assert removes / adds == pytest.approx(n / (1 + n), rel=2e-2)
        # the window never goes negative: removes never outrun adds
        assert np.all(np.cumsum(np.where(acts, 1, -1)) >= 0)


def test_accelerating_schedule_is_piecewise_loglinear():
    s = infer.AcceleratingSchedule(500.0, 4e3, 1e9)
    assert s.extra_passes(10) == 500.0 and s.extra_passes(4000) == 500.0
    assert s.extra_passes(2e9) == 0.0
    mid = s.extra_passes(2e6)       # log-midpoint: sqrt(501) passes
    assert mid + 1 == pytest.approx(501 ** 0.5, rel=1e-9)
    assert s.extra_passes(0.9e9) == 0.0        # would be barely above one pass: cut to one
    xs = np.logspace(np.log10(4e3), 9, 50)
    ps = [s.extra_passes(x) for x in xs]
    assert all(a >= b for a, b in zip(ps, ps[1:]))
    with pytest.raises(ValueError):
        infer.AcceleratingSchedule(1.0, 10, 5)


def test_batching_fires_when_every_assigned_row_was_replaced():
    b = infer.BatchingSchedule(0)
    b.add()
    assert b.test() and b.stale == 1 and b.fresh == 0     # first row: nothing stale
    b.add(); b.add()
    assert not b.test()
    b.remove()
    assert b.test() and b.stale == 2


def test_disabling_schedule():
    d = infer.KernelDisablingSchedule(2)
    for _ in range(2):
        d.run(False)
    assert d.test()
    d.run(False)
    assert not d.test()
    d.run(True)
    assert d.test()


def test_action_stream_is_a_moving_window():
    n_rows = 300
    acc = infer.AcceleratingSchedule(4.0, 50, 1e9)
    ann = infer.AnnealingSchedule(acc.extra_passes(0))
    bat, win = infer.BatchingSchedule(0), infer.RowWindow(n_rows)
    assigned, live, total, batches = 0, [], 0, 0
    while assigned != n_rows:
        acts, assigned, boundary = infer.next_batch(n_rows, assigned, ann, bat, win)
        for a in acts:
            row = int(a >> 1)
            if a & 1:
                assert live and live[0] == row        # REMOVE takes the oldest assigned row (FIFO)
                live.pop(0)
            else:
                assert row not in live                # never a duplicate row
                live.append(row)
        total += len(acts)
        assert len(live) == assigned
        if boundary:
            batches += 1
            ann.set_extra_passes(acc.extra_passes(assigned))
    assert sorted(live) == list(range(n_rows))
    assert batches > 5 and total > 2 * n_rows         # the extra passes really happened


def _run(abi, n_chains, kind_iterations):
    model, rows, _ = synth.generate(240, [("bb", 3), ("dd4", 3), ("gp", 2), ("nich", 2)], density=0.8, seed=12)
    engines = []
    for c in range(n_chains):
        e = Engine(abi)
        e.set_model(model)
        if c == 0:
            e.set_rows(rows)
        else:
            e.share_rows(engines[0])
        engines.append(e)
    cfg = {"schedule": {"extra_passes": 3.0, "small_data_size": 60, "big_data_size": 1e5, "max_reject_iters": 3},
           "kernels": {"kind": {"iterations": kind_iterations, "empty_kind_count": 4}}}
    grids = {"topology": [(2.0, 0.1), (10.0, 0.1)], "clustering": [(2.0, 0.1), (10.0, 0.1)],
             "bb": {"alpha": [0.5, 2.0], "beta": [0.5, 2.0]}, "dd": {"alpha": [0.5, 1.5]},
             "gp": {"alpha": [0.5, 1.5], "inv_beta": [0.5, 1.5]},
             "nich": {"kappa": [0.5, 1.5], "mu": [-1.0, 1.0], "nu": [0.5, 1.5], "sigmasq": [0.5, 1.5]}}
    logs = []
    batches = infer.infer_multi_pass(engines, list(range(5, 5 + n_chains)), cfg, hyper_prior=grids, log=logs.append)
    return model, rows, engines, batches, logs


def _check(model, rows, engines, batches, logs, kinds_may_move):
    assert batches == len(logs) and batches >= 3
    assert [l["assigned"] for l in logs] == sorted(l["assigned"] for l in logs)
    for e in engines:
        f2k = e.get_kinds()[1]
        K = e.n_kinds()
        assert len(f2k) == model.n_features and f2k.max() < K
        for k in range(K):
            a = e.get_assignments(k)
            assert (a != 0xFFFFFFFF).all()                       # every row assigned at the end
            assert e.kind_info(k).sample_size == rows.n_rows
            assert e.kind_info(k).n_features > 0                 # ephemeral kinds are gone
        assert np.isfinite(e.score_data())
        if not kinds_may_move:
            assert np.array_equal(f2k, model.f2k)


def test_infer_multi_pass_cat_structure_on_the_oracle(oracle_abi):
    out = _run(oracle_abi, 2, 0)
    _check(*out, kinds_may_move=False)
    a, b = out[2]
    assert not np.array_equal(a.get_assignments(0), b.get_assignments(0))   # chains differ by seed


def test_infer_multi_pass_kind_structure_on_the_oracle(oracle_abi):
    _check(*_run(oracle_abi, 2, 4), kinds_may_move=True)


def test_default_config_and_grids_are_wired(oracle_abi):
    cfg = infer.fill_in_defaults({"schedule": {"extra_passes": 1.0}})
    assert cfg["schedule"]["extra_passes"] == 1.0 and cfg["schedule"]["small_data_size"] == 4e3
    assert cfg["kernels"]["kind"]["empty_kind_count"] == 32
    model, rows, _ = synth.generate(60, [("bb", 2), ("dd4", 1), ("gp", 1), ("nich", 1)], density=0.9, seed=2)
    e = Engine(oracle_abi)
    e.set_model(model)
    e.set_rows(rows)
    infer.infer_single_pass([e], [3])
    e.set_hyper_prior(hyperprior.defaults())      # loom's full default grids
    e.hyper_run(seed=4)
    assert np.isfinite(e.score_data())


@pytest.mark.gpu
def test_infer_multi_pass_on_the_gpu(cuda_abi):
    _check(*_run(cuda_abi, 3, 0), kinds_may_move=False)
    _check(*_run(cuda_abi, 2, 4), kinds_may_move=True)


def test_schedules_reproduce_the_reference_binary():
    """src/schedules.hpp, compiled in place in the build container (oracle/Makefile `ref`) and driven through loom.cc's loop; its ADD / REMOVE decisions and batch boundaries
    are the committed fixture (tests/golden/make_reference_schedules.py).  The product's schedule driver must make
    exactly the same decisions: this pins WHEN every kernel fires against the reference itself."""
    import json
    import os
    from loom_b200 import infer as driver
    fixture = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_schedules.json")))
    assert len(fixture["cases"]) >= 10
    for case in fixture["cases"]:
        R = case["rows"]
        accelerating = driver.AcceleratingSchedule(case["extra_passes"], case["small"], case["big"])
        annealing = driver.AnnealingSchedule(case["extra_passes"])
        annealing.set_extra_passes(accelerating.extra_passes(0))
        batching, window = driver.BatchingSchedule(0), driver.RowWindow(R)
        limit = case["max_actions"] or float("inf")
        assigned, text, n = 0, [], 0
        while assigned != R and n < limit:
            acts, assigned, boundary = driver.next_batch(R, assigned, annealing, batching, window)
            text.append("".join("R" if a & 1 else "A" for a in acts))
            n += len(acts)
            if boundary:
                text.append("|")
                annealing.set_extra_passes(accelerating.extra_passes(assigned))
        got = "".join(text)
        want = case["decisions"]
        if case["assigned"] != R:                    # the reference run was cut by max_actions: compare the common prefix
            k = min(len(got), len(want))
            assert k > 5000 and got[:k] == want[:k], (case["extra_passes"], case["small"], case["big"], R)
        else:
            assert got == want, (case["extra_passes"], case["small"], case["big"], R)
            assert assigned == case["assigned"] == R


def test_cpp_host_schedules_reproduce_the_reference_binary(tmp_path):
    """Same fixture, the C++ host: the schedule classes of loom_b200/host/loom_b200.hpp (what loom_infer runs), compiled
    into tests/cpp/host_schedules_main.cc's loop, must print the reference's decision string."""
    import json
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "host_schedules")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(here, "cpp", "host_schedules_main.cc")])
    fixture = json.load(open(os.path.join(here, "golden", "reference_schedules.json")))
    for case in fixture["cases"]:
        args = [repr(case["extra_passes"]), repr(case["small"]), repr(case["big"]), str(case["rows"])]
        if case["max_actions"]:
            args.append(str(case["max_actions"]))
        decisions, assigned = subprocess.check_output([exe] + args).decode().split()
        assert decisions == case["decisions"], args
        assert int(assigned) == case["assigned"]
