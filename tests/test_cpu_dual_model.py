"""not gpu: a numpy model of the column side of the dual-direction pass --
thresholds from a strided row sample, row segments from B200._fused_segments, emits below the
threshold, compaction + threshold tightening between segments (kb2_col_compact), sticky overflow,
final selection with the bound the proof uses (kb2_col_select).  It pins the two invariants the
CUDA path relies on: (1) the cap best rows of every non-overflowed column survive, whatever the
segmentation; (2) every row that was not kept has key >= col_tau."""
import numpy as np
import pytest

from kiez_b200.neighbors import B200Mixin


class _Knobs:
    FUSED_SAMPLE_DIV = 32.0
    FUSED_SEGMENT_GROWTH = 3.0
    FUSED_SEGMENT_MIN_ROWS = 16
    FUSED_COL_CAP = 512


def column_side_model(keys, cap, knobs):
    """keys[row, col] = column key of (row, col).  Returns (kept rows per column or None where the
    column overflowed, col_tau, emits per column)."""
    n_rows, n_cols = keys.shape
    n_s = B200Mixin._fused_sample_rows(knobs, n_rows, cap)
    step = max(1, n_rows // n_s)
    sample = np.arange(n_s) * step
    tau = np.sort(keys[sample], axis=0)[cap - 1].copy()        # cap-th best within the sample
    col_cap = B200Mixin._fused_col_cap(knobs, cap)
    bufs = [[] for _ in range(n_cols)]                          # (key, row) entries per column
    overflowed = np.zeros(n_cols, dtype=bool)
    emits = np.zeros(n_cols, dtype=np.int64)
    bounds = B200Mixin._fused_segments(knobs, n_rows, n_s)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        for c in range(n_cols):
            rows = lo + np.flatnonzero(keys[lo:hi, c] < tau[c])
            emits[c] += rows.size
            if overflowed[c]:
                continue
            bufs[c].extend((keys[r, c], r) for r in rows)
            if len(bufs[c]) > col_cap:
                overflowed[c] = True                             # sticky: the column is re-searched
        if hi < n_rows:                                          # kb2_col_compact
            for c in range(n_cols):
                if not overflowed[c] and len(bufs[c]) >= cap:
                    bufs[c] = sorted(bufs[c])[:cap]
                    tau[c] = bufs[c][cap - 1][0]
    kept, col_tau = [], np.empty(n_cols)
    for c in range(n_cols):                                      # kb2_col_select
        entries = sorted(bufs[c])
        kept.append(None if overflowed[c] else [r for _k, r in entries[:cap]])
        col_tau[c] = entries[cap - 1][0] if len(entries) >= cap else tau[c]
    return kept, col_tau, emits, bounds


@pytest.mark.parametrize(("n_rows", "n_cols", "cap", "growth", "col_cap"), [
    (3000, 40, 16, 3.0, 512), (3000, 40, 16, 2.0, 512), (2500, 30, 8, 1.5, 512),
    (3000, 40, 16, 1.0, 4096), (3000, 40, 16, 3.0, 40)])
def test_column_side_model_keeps_the_exact_top_cap(n_rows, n_cols, cap, growth, col_cap):
    rng = np.random.default_rng(int(n_rows * growth) + cap)
    keys = rng.standard_normal((n_rows, n_cols)) + rng.standard_normal((1, n_cols))
    knobs = _Knobs()
    knobs.FUSED_SEGMENT_GROWTH, knobs.FUSED_COL_CAP = growth, col_cap
    kept, col_tau, emits, bounds = column_side_model(keys, cap, knobs)
    order = np.argsort(keys, axis=0, kind="stable")
    n_over = 0
    for c in range(n_cols):
        if kept[c] is None:
            n_over += 1
            continue
        assert kept[c] == list(order[:cap, c]), f"column {c}: lost one of its {cap} best rows"
        others = np.setdiff1d(np.arange(n_rows), kept[c])
        assert (keys[others, c] >= col_tau[c]).all(), f"column {c}: col_tau is not a lower bound"
    if col_cap == 40:
        assert n_over > 0                                       # the overflow path is exercised
    else:
        assert n_over == 0
    if growth > 1.0 and col_cap == 512:
        # tightening thresholds: ~cap (g - 1) emits per segment instead of cap * D from one bound
        n_s = B200Mixin._fused_sample_rows(knobs, n_rows, cap)
        assert len(bounds) > 2
        assert emits.mean() < 0.5 * cap * n_rows / n_s
