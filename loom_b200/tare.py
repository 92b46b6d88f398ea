"""Tare / sparsify: a restatement of loom's Differ for the columnar row block.

`make_tare` follows Differ::add_rows + _make_tare_type (src/differ.cc:93-206,
src/differ.hpp:63-95): a column's tare value is its mode when the mode covers more
than half of ALL rows; booleans and counts only (counts: mode searched in [0,16)),
reals are never tared.  `sparsify` follows Differ::_abs_to_rel_type
(src/differ.cc:248-298): for a tared column, a row that agrees with the tare stores
nothing, one that differs stores (pos = its value, neg = the tare cell), one that lacks
the cell stores neg; untared columns go to pos when observed.
Worked example: doc/adapting.md:323-414 (tests/test_tare.py).
"""
import numpy as np

MAX_COUNT = 16   # CountSummary::max_count (src/differ.hpp:66)


def make_tare(model, rows):
    """-> (tare_observed uint8 [F], tare_values uint32 [F])"""
    F, R = model.n_features, rows.n_rows
    obs = rows.observed()
    F8 = model.n8()
    tare_obs = np.zeros(F, np.uint8)
    tare_val = np.zeros(F, np.uint32)
    for f, feat in enumerate(model.features):
        if feat["type"] == "nich":
            continue                      # "do not sparsify reals" (differ.cc:121)
        col = rows.cat8[f] if f < F8 else rows.wide[f - F8]
        vals = col[obs[f]].astype(np.int64)
        limit = 2 if feat["type"] == "bb" else MAX_COUNT
        counts = np.bincount(vals[vals < limit], minlength=limit)[:limit]
        mode = int(np.argmax(counts))     # first maximum, like get_mode
        if counts[mode] > 0.5 * R:
            tare_obs[f] = 1
            tare_val[f] = mode
    return tare_obs, tare_val


def sparsify(model, rows, tare_obs, tare_val):
    """-> dict of the arrays lb_set_rows_diff takes (every row references the tare)."""
    F, R = model.n_features, rows.n_rows
    F8 = model.n8()
    obs = rows.observed()
    vals = np.zeros((F, R), np.uint32)
    vals[:F8] = rows.cat8
    vals[F8:] = rows.wide
    t_on = tare_obs.astype(bool)[:, None]
    differs = obs & (vals != tare_val[:, None])
    pos = np.where(t_on, differs, obs)                  # [F][R]
    neg = t_on & (~obs | differs)
    def csr(m):
        mt = m.T                                        # [R][F]: entries of a row in feature order
        ptr = np.zeros(R + 1, np.uint64)
        ptr[1:] = np.cumsum(mt.sum(axis=1))
        r_idx, f_idx = np.nonzero(mt)
        return ptr, f_idx.astype(np.uint32), r_idx
    pos_ptr, pos_f, pos_r = csr(pos)
    neg_ptr, neg_f, _ = csr(neg)
    return dict(n_rows=R, tare_observed=tare_obs, tare_values=tare_val, row_has_tare=None,
                pos_ptr=pos_ptr, pos_feature=pos_f, pos_value=vals[pos_f, pos_r].astype(np.uint32),
                neg_ptr=neg_ptr, neg_feature=neg_f)
