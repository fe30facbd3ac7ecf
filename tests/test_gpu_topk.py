"""Row-wise top-k kernels behind HubnessReduction._sort (kiez/hubness_reduction/base.py:72-87) and
the multi-GPU merge, through the C ABI, over every code path of csrc/rescale.cu: one thread per
row (width <= 16), the warp arg-min selection (width > 16, k <= 16), the bitonic sort over lanes x
registers (k > 16) and the shared-memory sort (width > 256).  Order: ascending value, ties by
input position, NaN last -- what numpy's stable argsort gives."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _want(dist, ind, k):
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]       # NaN last, ties by position
    return np.take_along_axis(dist, order, 1), np.take_along_axis(ind, order, 1)


def _topk(dist, ind, k, nparts=1):
    from kiez_b200 import _lib

    dev = torch.device("cuda", 0)
    n, c = dist.shape[-2], dist.shape[-1]
    d_t = torch.from_numpy(np.ascontiguousarray(dist)).to(dev)
    i_t = torch.from_numpy(np.ascontiguousarray(ind)).to(dev)
    out_d = torch.full((n, k), -1.0, dtype=torch.float64, device=dev)
    out_i = torch.full((n, k), -7, dtype=torch.int64, device=dev)
    _lib.call("kb2_topk_rows", _lib.ptr(d_t), _lib.ptr(i_t), n, c, nparts, n * c if nparts > 1 else 0,
              k, _lib.ptr(out_d), _lib.ptr(out_i), _lib.stream_ptr())
    torch.cuda.synchronize()
    return out_d.cpu().numpy(), out_i.cpu().numpy()


@pytest.mark.parametrize(("c", "k"), [
    (3, 1), (8, 8), (10, 10), (10, 5), (12, 7), (16, 16),            # one thread per row
    (17, 3), (32, 10), (50, 10), (50, 16), (64, 1), (100, 10), (128, 5), (200, 10), (256, 16),  # arg-min rounds
    (17, 17), (50, 50), (100, 40), (256, 100),                         # bitonic sort over lanes x registers
    (300, 10), (300, 120),                                             # shared-memory sort
])
def test_topk_rows_matches_stable_argsort(c, k):
    rng = np.random.default_rng(1000 * c + k)
    n = 777                                                            # ragged: not a multiple of any tile
    dist = rng.standard_normal((n, c))
    # ties (rounded values), NaN runs, signed zeros, infinities
    dist[: n // 3] = np.round(dist[: n // 3], 1)
    dist[rng.random((n, c)) < 0.05] = np.nan
    dist[rng.random((n, c)) < 0.02] = 0.0
    dist[rng.random((n, c)) < 0.02] = -0.0
    dist[rng.random((n, c)) < 0.01] = np.inf
    dist[rng.random((n, c)) < 0.01] = -np.inf
    dist[5] = np.nan                                                   # a row of NaN only
    dist[6] = 1.25                                                     # a row of ties only
    ind = rng.integers(0, 10**9, size=(n, c))
    got_d, got_i = _topk(dist, ind, k)
    want_d, want_i = _want(dist, ind, k)
    np.testing.assert_array_equal(got_d, want_d)                      # NaN == NaN here; -0.0 == 0.0
    np.testing.assert_array_equal(got_i, want_i)


@pytest.mark.parametrize(("c", "nparts", "k"), [(10, 1, 10), (5, 3, 5), (10, 8, 10), (25, 4, 10), (50, 4, 50)])
def test_topk_rows_merges_parts(c, nparts, k):
    """nparts > 1: row r is the concatenation of the parts' rows (the multi-GPU merge)."""
    rng = np.random.default_rng(c * 31 + nparts)
    n = 1234
    dist = np.sort(rng.standard_normal((nparts, n, c)), axis=2)        # per-rank lists arrive sorted
    dist[:, : n // 4] = np.round(dist[:, : n // 4], 1)
    ind = rng.integers(0, 10**9, size=(nparts, n, c))
    got_d, got_i = _topk(dist, ind, k, nparts=nparts)
    cat_d = np.concatenate(list(dist), axis=1)
    cat_i = np.concatenate(list(ind), axis=1)
    want_d, want_i = _want(cat_d, cat_i, k)
    np.testing.assert_array_equal(got_d, want_d)
    np.testing.assert_array_equal(got_i, want_i)
