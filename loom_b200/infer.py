"""Host driver of a whole inference run over the device engine: loom's schedules and the
`Loom::infer_*` loops, restated.

    schedules      src/schedules.hpp:41-321   when rows are added / removed and when the batch kernels fire
    row window     src/stream_interval.hpp:36-140   two cyclic cursors over the shuffled rows
    driver loops   src/loom.cc:100-165 (single pass), :167-450 (multi pass, cat / kind structure)

The reference decides action by action and blocks at every batch boundary (`pipeline.wait()`).
Here the same decisions are made up front for a whole batch -- the schedules depend only on
counts, never on sampled values -- and the batch is handed to the device as ONE action list
(`lb_cat_sweep`, or `lb_cat_sweep_multi` for many chains that share the schedule).  At a batch
boundary the kind kernel, the hyper kernel and the schedule updates run in the reference's order.
This is host scalar logic; nothing here touches the row data.
"""
import math

import numpy as np

from .engine import cat_sweep_multi, cat_sweep_streams

DEFAULT_CONFIG = {   # loom/config.py:34-74
    "seed": 0,
    "schedule": {"extra_passes": 500.0, "small_data_size": 4e3, "big_data_size": 1e9, "max_reject_iters": 100},
    "kernels": {"cat": {"empty_group_count": 1},
                "hyper": {"run": True},
                "kind": {"iterations": 32, "empty_kind_count": 32}},
}


def fill_in_defaults(config, defaults=DEFAULT_CONFIG):
    out = dict(config or {})
    for key, value in defaults.items():
        if isinstance(value, dict):
            out[key] = fill_in_defaults(out.get(key), value)
        else:
            out.setdefault(key, value)
    return out


class AnnealingSchedule:
    """ADD / REMOVE interleaving with REMOVE : ADD = N : (1 + N) for N extra passes
    (src/schedules.hpp:41-108): a counter pays `N` per ADD and earns `1 + N` per REMOVE; the next
    action is an ADD while the counter is non-negative."""
    INIT_ROW_COUNT = 2
    MAX_EXTRA_PASSES = 1000000

    def __init__(self, extra_passes):
        self.set_extra_passes(extra_passes)
        self.state = self.INIT_ROW_COUNT * self.add_rate

    def set_extra_passes(self, extra_passes):
        if not 0 <= extra_passes <= self.MAX_EXTRA_PASSES:
            raise ValueError("extra_passes out of range: %r" % extra_passes)
        self.add_rate = 1.0 + extra_passes
        self.remove_rate = float(extra_passes)

    def next_action_is_add(self):
        if self.state >= 0:
            self.state -= self.remove_rate
            return True
        self.state += self.add_rate
        return False


class AcceleratingSchedule:
    """Fewer extra passes as the assigned window grows: constant up to `small_data_size`, then a
    straight line in log(rows)-log(passes) down to one pass at `big_data_size`
    (src/schedules.hpp:110-165); totals of at most 2 passes are cut to a single pass."""

    def __init__(self, extra_passes, small_data_size, big_data_size):
        if not 0 <= small_data_size <= big_data_size:
            raise ValueError("need 0 <= small_data_size <= big_data_size")
        if math.isinf(big_data_size) and not math.isinf(small_data_size):
            raise ValueError("big_data_size infinite but small_data_size finite")
        self.base, self.small, self.big = float(extra_passes), float(small_data_size), float(big_data_size)

    def extra_passes(self, row_count):
        if row_count <= self.small:
            return self.base
        if row_count > self.big:
            return 0.0
        if self.small == 0:           # the reference divides by log(big / 0) = inf: power 0, one pass (IEEE arithmetic in C++)
            return 0.0
        power = math.log(self.big / row_count) / math.log(self.big / self.small)
        passes = (1.0 + self.base) ** power
        return passes - 1.0 if passes > 2.0 else 0.0


class BatchingSchedule:
    """A batch ends when every row that was assigned at its start has been replaced
    (src/schedules.hpp:167-208)."""

    def __init__(self, row_count=0):
        self.stale, self.fresh = int(row_count), 0

    def add(self):
        self.fresh += 1

    def remove(self):
        self.stale -= 1

    def test(self):
        if self.stale == 0 and self.fresh:
            self.stale, self.fresh = self.fresh, 0
            return True
        return False


class KernelDisablingSchedule:
    """Switch the kind kernel off after `max_reject_iters` batches in a row without an accepted
    move (src/schedules.hpp:210-247)."""

    def __init__(self, max_reject_iters):
        self.max_reject_iters, self.reject_iters = int(max_reject_iters), 0

    def run(self, accepted):
        self.reject_iters = 0 if accepted else self.reject_iters + 1

    def test(self):
        return self.reject_iters <= self.max_reject_iters


class RowWindow:
    """StreamInterval (src/stream_interval.hpp:36-140): the assigned rows are a window of the
    cyclic shuffled stream; ADD takes the row at the front cursor, REMOVE the oldest assigned row
    at the back cursor.  Positions index the device row block."""

    def __init__(self, n_rows):
        self.n_rows, self.unassigned, self.assigned = int(n_rows), 0, 0

    def read_unassigned(self):
        pos = self.unassigned
        self.unassigned = (pos + 1) % self.n_rows
        return pos

    def read_assigned(self):
        pos = self.assigned
        self.assigned = (pos + 1) % self.n_rows
        return pos


def next_batch(n_rows, assigned_count, annealing, batching, window):
    """Actions up to and including the next batch boundary (or until every row is assigned).
    Returns (actions uint32, assigned_count, hit_boundary)."""
    actions = []
    while assigned_count != n_rows:
        if annealing.next_action_is_add():
            actions.append(window.read_unassigned() << 1)
            assigned_count += 1
            batching.add()
        else:
            actions.append((window.read_assigned() << 1) | 1)
            assigned_count -= 1
            batching.remove()
        if batching.test():
            return np.asarray(actions, np.uint32), assigned_count, True
    return np.asarray(actions, np.uint32), assigned_count, False


def sweep(engines, actions, seeds, action_offset=0, orders=None):
    """One batch of stream-position actions on every chain.  `orders` ([n_chains][n_rows], optional)
    is each chain's own shuffle of the shared row block (loom/tasks.py:228-233: loom.runner.shuffle
    with the sample's seed): stream position p of chain c is row orders[c][p]."""
    if orders is None:
        cat_sweep_multi(engines, actions, seeds, action_offset=action_offset)
        return
    orders = np.asarray(orders, np.uint32)
    per_chain = (orders[:, actions >> 1] << 1) | (actions & 1)[None, :]
    cat_sweep_streams(engines, per_chain, seeds, action_offset=action_offset)


def infer_single_pass(engines, seeds, orders=None):
    """Loom::infer_single_pass (src/loom.cc:100-165): one greedy ADD pass over the rows."""
    n_rows = engines[0].n_rows
    sweep(engines, (np.arange(n_rows, dtype=np.uint32) << 1), seeds, orders=orders)


def infer_multi_pass(engines, seeds, config=None, hyper_prior=None, log=None, orders=None):
    """Loom::infer_multi_pass (src/loom.cc:167-450) on one or many chains with a common schedule.

    engines      chains (models and rows already set; every chain sees the same rows)
    seeds        one per chain
    config       dict shaped like loom/config.py DEFAULTS (missing entries are filled in)
    hyper_prior  grids for the hyper kernel (loom_b200.hyperprior.defaults()); None = hyper kernel off
    log          optional callable(dict) called at every batch boundary (Logger of src/logger.hpp)
    orders       optional [n_chains][n_rows]: a row order per chain (each sample's own shuffle)

    The kind kernel runs while `kernels.kind.iterations` > 0 and the disabling schedule allows it,
    per chain (a chain whose proposals keep being rejected falls back to the cat-only loop, as
    infer_kind_structure returning false does, src/loom.cc:203-209).  Returns the batch count.
    """
    cfg = fill_in_defaults(config)
    sched = cfg["schedule"]
    n_rows = engines[0].n_rows
    accelerating = AcceleratingSchedule(sched["extra_passes"], sched["small_data_size"], sched["big_data_size"])
    annealing = AnnealingSchedule(accelerating.extra_passes(0))
    batching = BatchingSchedule(0)
    window = RowWindow(n_rows)
    kind_cfg = cfg["kernels"]["kind"]
    kind_on = [kind_cfg["iterations"] > 0] * len(engines)
    disabling = [KernelDisablingSchedule(sched["max_reject_iters"]) for _ in engines]
    run_hyper = bool(cfg["kernels"]["hyper"]["run"]) and hyper_prior is not None
    seeds = np.asarray(seeds, np.uint64)
    if run_hyper:
        for e in engines:
            e.set_hyper_prior(hyper_prior)
    for c, e in enumerate(engines):
        if kind_on[c]:
            e.init_featureless_kinds(kind_cfg["empty_kind_count"], int(seeds[c]) * 7919 + 1)
    assigned, offset, batches = 0, 0, 0
    while assigned != n_rows:
        actions, assigned, boundary = next_batch(n_rows, assigned, annealing, batching, window)
        if len(actions):
            sweep(engines, actions, seeds, action_offset=offset, orders=orders)
            offset += len(actions)
        if not boundary:
            break
        batches += 1
        annealing.set_extra_passes(accelerating.extra_passes(assigned))
        for c, e in enumerate(engines):
            s = int(seeds[c]) * 7919 + 1000 * batches
            if kind_on[c]:
                e.kind_proposer_score(fetch=False)
                _, changed = e.kind_run(kind_cfg["iterations"], s + 2)
                disabling[c].run(changed > 0)
                if disabling[c].test():
                    e.init_featureless_kinds(kind_cfg["empty_kind_count"], s + 3)
                else:            # kind structure frozen from here on: drop the ephemeral kinds
                    e.init_featureless_kinds(0, s + 3)
                    kind_on[c] = False
            if run_hyper:
                e.hyper_run(s + 4)
        if log:
            log({"iter": batches, "assigned": assigned, "actions": offset,
                 "extra_passes": annealing.remove_rate, "kind_kernel_on": sum(kind_on)})
    for c, e in enumerate(engines):
        if kind_on[c]:
            e.init_featureless_kinds(0, int(seeds[c]) * 7919 + 5)   # KindKernel::~KindKernel (kind_kernel.cc:75-81)
    return batches
