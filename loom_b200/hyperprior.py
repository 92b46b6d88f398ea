"""Default hyper-prior grids of the HyperKernel (loom/hyperprior.py:33-71 DEFAULTS over
loom/gridding.py:32-88), in the dict form Engine.set_hyper_prior takes.

Grid shapes: `uniform` = midpoints of equal cells; `center_heavy` = arcsine-spaced (dense at both
ends of the unit interval mapped through asin, i.e. dense in the middle of the value range);
`left_heavy` = squared uniform (dense near the minimum); `right_heavy` = its mirror image.
The Pitman-Yor grid keeps the lower triangle x + y < 1 of center_heavy(alpha) x left_heavy(d)
on the unit square, alpha log-spaced over [0.1, 100], d linear over [0, 0.5].
"""
import numpy as np


def uniform(lo, hi, n):
    return lo + (np.arange(n) + 0.5) * (hi - lo) / float(n)


def center_heavy(lo, hi, n):
    return lo + (np.arcsin(uniform(-1.0, 1.0, n)) / np.pi + 0.5) * (hi - lo)


def left_heavy(lo, hi, n):
    return lo + uniform(0.0, 1.0, n) ** 2 * (hi - lo)


def right_heavy(lo, hi, n):
    return left_heavy(hi, lo, n)[::-1].copy()


def pitman_yor(min_alpha=0.1, max_alpha=100.0, min_d=0.0, max_d=0.5, alpha_count=20, d_count=10):
    grid = []
    for x in center_heavy(0.0, 1.0, alpha_count):
        for y in left_heavy(0.0, 1.0, d_count):
            if x + y < 1.0:
                grid.append({"alpha": float(min_alpha * (max_alpha / min_alpha) ** x),
                             "d": float(min_d + (max_d - min_d) * y)})
    return grid


def defaults():
    dd_alpha = np.logspace(-1, 2, 12).tolist()
    pos = np.logspace(-8, 8, 100).tolist()
    neg = (-np.logspace(-8, 8, 100)).tolist()
    py = pitman_yor()
    return {
        "topology": py,
        "clustering": py,
        "bb": {"alpha": dd_alpha, "beta": dd_alpha},
        "dd": {"alpha": dd_alpha},
        "dpd": {"gamma": (10 ** left_heavy(-1, 2, 30)).tolist(), "alpha": (10 ** right_heavy(-1, 1, 20)).tolist()},
        "gp": {"alpha": np.logspace(-1, 5, 100).tolist(), "inv_beta": np.logspace(-5, 1, 100).tolist()},
        "nich": {"mu": neg + [0.0] + pos, "sigmasq": pos, "kappa": np.logspace(-2, 2, 30).tolist(),
                 "nu": np.logspace(0, 2, 30).tolist()},
    }
