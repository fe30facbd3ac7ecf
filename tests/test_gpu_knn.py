"""-m gpu: the candidate search + exact finish (B200._kneighbors) against the CPU oracle,
through the C-ABI, for both search kernels (tcgen05 and the FP32-pipe cross-check)."""
import numpy as np
import pytest

from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RTOL, ATOL = 1e-5, 5e-6   # north star: 1e-5 relative; atol covers sklearn's sqrt(noise) self-distances
                          # (expanded form: sqrt(eps * ||x||^2 * few) ~ 1e-7..1e-6 where the true distance is 0)


def _algo(**kw):
    from kiez_b200 import B200

    return B200(**kw)


def _data(nq, ny, d, seed=0, dist="gauss"):
    rng = np.random.default_rng(seed)
    if dist == "gauss":
        q = rng.standard_normal((nq, d))
        y = rng.standard_normal((ny, d))
    elif dist == "shifted":      # far from the origin: exercises the centring
        q = 50.0 + rng.standard_normal((nq, d))
        y = 50.0 + rng.standard_normal((ny, d))
    else:                        # clustered unit vectors: near ties
        cent = rng.standard_normal((16, d))
        q = cent[rng.integers(0, 16, nq)] + 0.05 * rng.standard_normal((nq, d))
        y = cent[rng.integers(0, 16, ny)] + 0.05 * rng.standard_normal((ny, d))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        y /= np.linalg.norm(y, axis=1, keepdims=True)
    return q.astype(np.float32), y.astype(np.float32)


SHAPES = [
    # nq, ny, d, k
    (1, 1, 1, 1),
    (3, 7, 5, 7),            # k == ny
    (20, 50, 5, 5),          # the reference's conftest shape
    (100, 100, 50, 10),      # README shape
    (127, 255, 31, 8),
    (129, 257, 33, 16),
    (300, 1000, 64, 50),
    (513, 2049, 128, 100),   # c = 100 (config C3's candidate count)
    (1000, 5000, 256, 10),
]


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt"])
@pytest.mark.parametrize(("nq", "ny", "d", "k"), SHAPES)
@pytest.mark.parametrize("metric", ["euclidean", "cosine", "sqeuclidean"])
def test_knn_matches_oracle(impl, nq, ny, d, k, metric):
    q, y = _data(nq, ny, d, seed=nq + ny)
    algo = _algo(n_candidates=k, metric=metric, impl=impl)
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=k)                      # forward: source rows vs target index
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), k, metric)
    assert dist.dtype == torch.float64 and ind.dtype == torch.int64 and dist.is_cuda
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"fwd {impl} {metric} {nq}x{ny}x{d}")
    k_rev = min(k, nq)
    dist, ind = algo.kneighbors(k=k_rev, query=y, s_to_t=False)   # reverse
    want_d, want_i = O.knn_brute(y.astype(np.float64), q.astype(np.float64), k_rev, metric)
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"rev {impl} {metric} {nq}x{ny}x{d}")


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt"])
@pytest.mark.parametrize("dist_kind", ["shifted", "hubby"])
def test_knn_hard_distributions(impl, dist_kind):
    q, y = _data(700, 1500, 96, seed=3, dist=dist_kind)
    algo = _algo(n_candidates=20, impl=impl)
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=20)
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 20, "euclidean")
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"{impl} {dist_kind}")


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt"])
def test_self_query_excludes_self(impl):
    """sklearn kneighbors(X=None) semantics (neighbors/_base.py:937-958)."""
    q, _ = _data(600, 1, 40, seed=5)
    algo = _algo(n_candidates=7, impl=impl)
    algo.fit(q)                                            # single-source mode
    dist, ind = algo.kneighbors(k=7)
    ind_np = ind.cpu().numpy()
    assert not (ind_np == np.arange(600)[:, None]).any()
    want_d, want_i = O.knn_brute(q.astype(np.float64), q.astype(np.float64), 7, "euclidean",
                                 exclude_self=True)
    O.assert_neighbors_match(dist.cpu().numpy(), ind_np, want_d, want_i, RTOL, ATOL, what="self")
    # explicit query = same data: self is NOT excluded (the reverse pass of HubnessReduction.fit)
    dist, ind = algo.kneighbors(k=3, query=q, s_to_t=False)
    assert (ind[:, 0].cpu().numpy() == np.arange(600)).all()
    assert float(dist[:, 0].abs().max()) == 0.0


def test_tc_and_simt_agree_with_many_splits():
    """Index splits (used to fill the GPU when there are few query tiles) + merge."""
    q, y = _data(200, 20000, 64, seed=9)
    outs = []
    for impl in ("tc", "tc1", "simt"):
        algo = _algo(n_candidates=10, impl=impl)
        qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
        for splits in (1, 3, 8):
            d, i = algo.search(qp, yp, 10, splits=splits)
            outs.append((d.cpu().numpy(), i.cpu().numpy()))
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 10, "euclidean")
    for d, i in outs:
        O.assert_neighbors_match(d, i, want_d, want_i, RTOL, ATOL, what="splits")


def test_return_types_and_torch_input():
    q, y = _data(64, 100, 16, seed=2)
    algo = _algo(n_candidates=5)
    algo.fit(torch.from_numpy(q).cuda(), torch.from_numpy(y).cuda())
    ind = algo.kneighbors(k=5, return_distance=False)
    d2, ind2 = algo.kneighbors(k=5, return_distance=True)
    assert torch.equal(ind, ind2)                          # tests/neighbors/test_sklearn.py:9-15
    assert (torch.diff(d2, dim=1) >= 0).all()
    with pytest.warns(UserWarning, match="larger than number of samples"):
        d3, _ = algo.kneighbors(k=500)
    assert d3.shape == (64, 100)                           # clamped to the index size


def test_large_sampled_rows():
    """Full-size-ish property check: every sampled row of a 20k x 30k search equals the oracle."""
    q, y = _data(20000, 30000, 64, seed=11)
    algo = _algo(n_candidates=10)
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=10)
    rows = np.random.default_rng(0).choice(20000, 256, replace=False)
    want_d, want_i = O.knn_brute(q[rows].astype(np.float64), y.astype(np.float64), 10, "euclidean")
    O.assert_neighbors_match(dist[rows].cpu().numpy(), ind[rows].cpu().numpy(), want_d, want_i,
                             RTOL, ATOL, what="sampled")
    d = dist.cpu().numpy()
    assert (np.diff(d, axis=1) >= 0).all()                 # sortedness at full size
    i = ind.cpu().numpy()
    assert ((i >= 0) & (i < 30000)).all()
    assert (np.sort(i, axis=1)[:, 1:] != np.sort(i, axis=1)[:, :-1]).all()   # no duplicate ids


@pytest.mark.parametrize(("nq", "ny", "d", "c"), [(300, 500, 32, 5), (1000, 1300, 64, 10),
                                                   (2049, 4097, 256, 10), (5000, 700, 128, 50),
                                                   (700, 5000, 96, 24)])
@pytest.mark.parametrize("single", [False, True])
def test_fused_dual_direction_pass(nq, ny, d, c, single):
    """One contraction, both directions (B200.search_both): row-wise results must equal the
    forward search, column-wise results the reverse search, both equal to the oracle."""
    q, y = _data(nq, ny, d, seed=nq + d)
    if single:
        y = q
        ny = nq
    algo = _algo(n_candidates=c, fused=True)
    qp = algo._prepare(q, cache=False)
    yp = qp if single else algo._prepare(y, cache=False)
    k_fwd = min(c, ny - (1 if single else 0))
    k_rev = min(c, nq)
    (fd, fi), (rd, ri) = algo.search_both(qp, yp, k_fwd, k_rev, exclude_self_rows=single)
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    want_d, want_i = O.knn_brute(q64, y64, k_fwd, "euclidean", exclude_self=single)
    O.assert_neighbors_match(fd.cpu().numpy(), fi.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="fwd")
    want_d, want_i = O.knn_brute(y64, q64, k_rev, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")


def test_fused_column_overflow_falls_back():
    """A column buffer that overflows (here: forced by a tiny capacity) is re-searched."""
    q, y = _data(3000, 400, 32, seed=8)
    algo = _algo(n_candidates=10, fused=True)
    algo.FUSED_COL_CAP = 16                  # == cap: every column with > 16 emitted rows overflows
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    (_fd, _fi), (rd, ri) = algo.search_both(qp, yp, 10, 10)
    want_d, want_i = O.knn_brute(y.astype(np.float64), q.astype(np.float64), 10, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")
