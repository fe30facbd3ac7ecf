"""CPU model of the wide-row top-k selection of the rescale kernels (csrc/rescale.cu,
`sortable_bits` + `warp_select_topk`): values are mapped to order-preserving 64-bit integers
(NaN last, -0.0 folded onto +0.0), and k rounds of a warp-wide arg-min -- minimum of the high
words, then of the low words among the lanes that match, then of the positions -- pick the k
smallest (value, position) pairs in order.  Restated here with numpy and checked against numpy's
stable argsort, the order HubnessReduction._sort relies on (kiez/hubness_reduction/base.py:72-87)."""
import numpy as np
import pytest

FULL = (1 << 64) - 1


def sortable_bits(v):
    """rescale.cu sortable_bits: unsigned order == (value ascending, NaN last)."""
    if np.isnan(v):
        return FULL
    v = v + 0.0                                            # -0.0 -> +0.0
    u = int(np.array(v, dtype=np.float64).view(np.uint64))
    return (~u & FULL) if (u >> 63) else (u | (1 << 63))


def select_topk(row, k, lanes=32):
    """warp_select_topk: element e lives in register e // lanes of lane e % lanes."""
    total = len(row)
    regs = (total + lanes - 1) // lanes
    u = {(ln, t): sortable_bits(row[t * lanes + ln]) if t * lanes + ln < total else FULL
         for ln in range(lanes) for t in range(regs)}
    pos = {(ln, t): t * lanes + ln if t * lanes + ln < total else 0xFFFFFFFF
           for ln in range(lanes) for t in range(regs)}
    out = []
    for _ in range(k):
        best = {}
        for ln in range(lanes):                            # lane-local best by (u, pos)
            best[ln] = min((u[ln, t], pos[ln, t]) for t in range(regs))
        mh = min(b[0] >> 32 for b in best.values())        # redux.min on the high words
        ml = min((b[0] & 0xFFFFFFFF) if (b[0] >> 32) == mh else 0xFFFFFFFF for b in best.values())
        match = {ln: (b[0] >> 32) == mh and (b[0] & 0xFFFFFFFF) == ml for ln, b in best.items()}
        mp = min(best[ln][1] if match[ln] else 0xFFFFFFFF for ln in range(lanes))
        winners = [ln for ln in range(lanes) if match[ln] and best[ln][1] == mp and mp != 0xFFFFFFFF]
        assert len(winners) == 1
        ln = winners[0]
        for t in range(regs):
            if pos[ln, t] == mp:
                out.append(mp)
                u[ln, t], pos[ln, t] = FULL, 0xFFFFFFFF    # retired
    return out


@pytest.mark.parametrize(("c", "k"), [(17, 3), (32, 10), (50, 10), (64, 16), (100, 10), (200, 16), (256, 7)])
def test_argmin_rounds_equal_stable_argsort(c, k):
    rng = np.random.default_rng(c * 100 + k)
    for trial in range(25):
        row = rng.standard_normal(c)
        if trial % 2:
            row = np.round(row, 1)                         # ties
        row[rng.random(c) < 0.1] = np.nan
        row[rng.random(c) < 0.05] = 0.0
        row[rng.random(c) < 0.05] = -0.0
        row[rng.random(c) < 0.03] = np.inf
        row[rng.random(c) < 0.03] = -np.inf
        if trial == 3:
            row[:] = np.nan
        want = list(np.argsort(row, kind="stable")[:k])
        assert select_topk(row, k) == want


def test_sortable_bits_is_order_preserving():
    vals = [-np.inf, -1e300, -1.5, -5e-324, -0.0, 0.0, 5e-324, 1.5, 1e300, np.inf, np.nan]
    bits = [sortable_bits(v) for v in vals]
    assert bits[4] == bits[5]                              # signed zeros tie
    assert all(a <= b for a, b in zip(bits, bits[1:]))
    assert all(a < b for a, b in zip(bits[:4], bits[1:5])) and all(a < b for a, b in zip(bits[5:], bits[6:]))
