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
    """impl="screen" = the tcgen05 pair kernels with the 1xTF32 screen + proof + 3xTF32
    re-search; every other impl pins the 3xTF32 keys so that both paths stay covered."""
    from kiez_b200 import B200

    if kw.get("impl") == "screen":
        kw.update(impl="tc", precision="screen")
    else:
        kw.setdefault("precision", "tf32x3")
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
    (600, 3000, 256, 50),    # c = 50 at d = 256 (C2 / C4-c50): partially resident query tile
    (400, 2500, 256, 100),   # c = 100 at d = 256 (C3): 2 of 8 K chunks resident
    (300, 900, 600, 10),     # dpad = 608 > 256: most of the query tile is streamed
]


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt", "screen"])
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


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt", "screen"])
@pytest.mark.parametrize("dist_kind", ["shifted", "hubby"])
def test_knn_hard_distributions(impl, dist_kind):
    q, y = _data(700, 1500, 96, seed=3, dist=dist_kind)
    algo = _algo(n_candidates=20, impl=impl)
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=20)
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 20, "euclidean")
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"{impl} {dist_kind}")


@pytest.mark.parametrize("impl", ["tc", "tc1", "simt", "screen"])
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
@pytest.mark.parametrize("impl", ["tc", "screen"])
def test_fused_dual_direction_pass(nq, ny, d, c, single, impl):
    """One contraction, both directions (B200.search_both): row-wise results must equal the
    forward search, column-wise results the reverse search, both equal to the oracle."""
    q, y = _data(nq, ny, d, seed=nq + d)
    if single:
        y = q
        ny = nq
    algo = _algo(n_candidates=c, fused=True, impl=impl)
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


@pytest.mark.parametrize("impl", ["tc", "screen"])
def test_fused_column_overflow_falls_back(impl):
    """A column buffer that overflows (here: forced by a tiny capacity) is re-searched."""
    q, y = _data(3000, 400, 32, seed=8)
    algo = _algo(n_candidates=10, fused=True, impl=impl)
    algo.FUSED_COL_CAP = 16                  # == cap: every column with > 16 emitted rows overflows
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    (_fd, _fi), (rd, ri) = algo.search_both(qp, yp, 10, 10)
    want_d, want_i = O.knn_brute(y.astype(np.float64), q.astype(np.float64), 10, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")


@pytest.mark.parametrize("impl", ["tc", "screen"])
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("col_cap", [512, 24])
def test_fused_row_segments_tighten_thresholds(impl, single, col_cap):
    """The dual-direction pass in row segments with kb2_col_compact between them (thresholds
    tighten, buffers are compacted; col_cap = 24 forces sticky overflows across segments):
    same results as the oracle in both directions."""
    nq, ny, d, c = 5000, 1500, 64, 10
    q, y = _data(nq, ny, d, seed=77)
    if single:
        y, ny = q, nq
    algo = _algo(n_candidates=c, fused=True, impl=impl)
    algo.FUSED_SEGMENT_MIN_ROWS = 256
    algo.FUSED_SEGMENT_GROWTH = 2.0
    algo.FUSED_SAMPLE_DIV = 32.0
    algo.FUSED_COL_CAP = col_cap
    algo._collect_stats = True
    qp = algo._prepare(q, cache=False)
    yp = qp if single else algo._prepare(y, cache=False)
    k_fwd = c
    (fd, fi), (rd, ri) = algo.search_both(qp, yp, k_fwd, c, exclude_self_rows=single)
    stats = algo._fused_stats
    assert len(stats["row_segments"]) >= 5, stats
    if col_cap == 512:
        assert stats["overflow_columns"] == 0
        # one threshold from the sample alone would emit ~cap * 32 = 512 rows per column
        assert stats["emitted_per_column_mean"] < 200, stats
    else:
        assert stats["overflow_columns"] > 0
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    want_d, want_i = O.knn_brute(q64, y64, k_fwd, "euclidean", exclude_self=single)
    O.assert_neighbors_match(fd.cpu().numpy(), fi.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="fwd")
    want_d, want_i = O.knn_brute(y64, q64, c, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")


# ---------------------------------------------------------------------------
# the 1xTF32 screen: proposals + float64 completeness proof + 3xTF32 re-search
# ---------------------------------------------------------------------------
def test_screen_is_used_and_mostly_proven():
    q, y = _data(3000, 8000, 256, seed=21)
    algo = _algo(n_candidates=10, impl="screen")
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=10)
    st = dict(algo.search_stats)
    assert st["screen_rows"] == 3000                       # the screen ran (no silent 3xTF32 path)
    assert st["screen_unverified"] < 300                   # and its proof carried almost every row
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 10, "euclidean")
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what="screen")


@pytest.mark.parametrize("single", [False, True])
def test_screen_unproven_rows_are_searched_again(single):
    """With an absurdly pessimistic error bound no row can be proven: every row must take the
    3xTF32 re-search (self-exclusion included) and still equal the oracle."""
    q, y = _data(900, 1200, 48, seed=22)
    algo = _algo(n_candidates=8, impl="screen")
    algo._eps_acc = lambda dpad: 0.9
    if single:
        algo.fit(q)
        y = q
    else:
        algo.fit(q, y)
    dist, ind = algo.kneighbors(k=8)
    assert algo.search_stats["screen_unverified"] == algo.search_stats["screen_rows"] == 900
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 8, "euclidean",
                                 exclude_self=single)
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what="re-search")


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_screen_near_ties_need_the_proof(metric):
    """Tight clusters: neighbours closer together than the TF32 rounding of the screen keys.
    The proof must flag those rows (not silently accept a wrong list)."""
    q, y = _data(1500, 4000, 64, seed=23, dist="hubby")
    algo = _algo(n_candidates=10, metric=metric, impl="screen")
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=10)
    assert algo.search_stats["screen_rows"] == 1500
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 10, metric)
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"near ties {metric}")


@pytest.mark.parametrize("fused", [False, True])
def test_screen_chained_ranges(monkeypatch, fused):
    """Enough query tiles for the L2-blocked order: the index is cut into ranges that are
    searched in order per query tile, the lists carried from range to range."""
    monkeypatch.setenv("KB2_SCREEN_RANGE_MB", "0.25")      # 1024-row ranges at d = 64
    nq, ny = 40000, 6100
    q, y = _data(nq, ny, 64, seed=24)
    algo = _algo(n_candidates=10, impl="screen", fused=fused)
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    steps, chained = algo._screen_plan(nq, ny, 64, 16)
    assert chained == 1 and steps == 6
    rows = np.random.default_rng(1).choice(nq, 300, replace=False)
    if fused:
        (fd, fi), (rd, ri) = algo.search_both(qp, yp, 10, 10)
        cols = np.random.default_rng(2).choice(ny, 200, replace=False)
        want_d, want_i = O.knn_brute(y[cols].astype(np.float64), q.astype(np.float64), 10, "euclidean")
        O.assert_neighbors_match(rd[cols].cpu().numpy(), ri[cols].cpu().numpy(), want_d, want_i,
                                 RTOL, ATOL, what="chained rev")
    else:
        fd, fi = algo.search(qp, yp, 10)
    want_d, want_i = O.knn_brute(q[rows].astype(np.float64), y.astype(np.float64), 10, "euclidean")
    O.assert_neighbors_match(fd[rows].cpu().numpy(), fi[rows].cpu().numpy(), want_d, want_i,
                             RTOL, ATOL, what="chained fwd")
    assert algo.search_stats["screen_unverified"] < 0.05 * algo.search_stats["screen_rows"]


def test_screen_not_taken_for_float64_but_for_wide_rows():
    """float64 callers keep the 3xTF32 search (the proof assumes fp32 operands).  Rows wider than
    the 256 features a resident query tile can hold take the screen with a partially resident
    tile; beyond dpad = 1024 the 3xTF32 search takes over."""
    rng = np.random.default_rng(3)
    algo = _algo(n_candidates=5, impl="screen")
    algo.fit(rng.standard_normal((200, 20)), rng.standard_normal((300, 20)))   # float64
    algo.kneighbors(k=5)
    assert algo.search_stats["screen_rows"] == 0
    for d, screened in ((300, 200), (1100, 0)):
        q, y = _data(200, 400, d, seed=4)
        algo = _algo(n_candidates=5, impl="screen")
        algo.fit(q, y)
        dist, ind = algo.kneighbors(k=5)
        assert algo.search_stats["screen_rows"] == screened
        want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 5, "euclidean")
        O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                                 what=f"wide rows d={d}")


def test_screen_partial_residency_plan():
    """kb2_screen_config: how the kernel divides shared memory between the resident part of the
    query tile and the ring, for the BASELINE.json shapes."""
    import ctypes as C

    from kiez_b200 import _lib

    def plan(dpad, cap, dual, ny=0):
        slots, res = C.c_int(0), C.c_int(0)
        stages = _lib.lib.kb2_screen_config(dpad, cap, dual, 0, ny, C.addressof(slots), C.addressof(res))
        return stages, slots.value, res.value

    assert plan(256, 16, 1)[2] == 8 and plan(256, 16, 1)[0] >= 4       # C4: fully resident
    st, _sl, res = plan(256, 56, 0)                                     # C2 / C4 at c = 50
    assert st >= 4 and 4 <= res < 8
    st, _sl, res = plan(256, 112, 0)                                    # C3
    assert st >= 4 and 1 <= res < 8
    assert plan(128, 56, 0)[2] == 4                                     # C5: fully resident
    assert plan(2048, 16, 0)[0] == 0 and plan(256, 136, 0)[0] == 0
    # short indexes with long lists: longer append buffers (fewer merges), less residency
    assert plan(256, 112, 0, 100_000)[1] > plan(256, 112, 0)[1] and plan(256, 112, 0, 100_000)[0] >= 4
    assert plan(256, 56, 0, 15_000)[1] > plan(256, 56, 0)[1]
    assert plan(256, 16, 1, 1_000_000) == plan(256, 16, 1)                # C4 is unaffected


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("resident", [None, 0, 3])
def test_screen_streamed_query_chunks(monkeypatch, fused, resident):
    """The ring carries query chunks next to the index tiles (KB2_SCREEN_RESIDENT forces fewer
    resident chunks than fit): same results as the oracle, chained ranges included."""
    if resident is not None:
        monkeypatch.setenv("KB2_SCREEN_RESIDENT", str(resident))
    monkeypatch.setenv("KB2_SCREEN_RANGE_MB", "1")
    nq, ny, d, c = 39000, 5000, 256, 50
    q, y = _data(nq, ny, d, seed=41)
    algo = _algo(n_candidates=c, impl="screen", fused=fused)
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    rows = np.random.default_rng(1).choice(nq, 200, replace=False)
    if fused:
        (fd, fi), (rd, ri) = algo.search_both(qp, yp, c, c)
        cols = np.random.default_rng(2).choice(ny, 100, replace=False)
        want_d, want_i = O.knn_brute(y[cols].astype(np.float64), q.astype(np.float64), c, "euclidean")
        O.assert_neighbors_match(rd[cols].cpu().numpy(), ri[cols].cpu().numpy(), want_d, want_i,
                                 RTOL, ATOL, what="streamed rev")
    else:
        fd, fi = algo.search(qp, yp, c)
    assert algo.search_stats["screen_rows"] == (nq + ny if fused else nq)
    want_d, want_i = O.knn_brute(q[rows].astype(np.float64), y.astype(np.float64), c, "euclidean")
    O.assert_neighbors_match(fd[rows].cpu().numpy(), fi[rows].cpu().numpy(), want_d, want_i,
                             RTOL, ATOL, what="streamed fwd")


def _tight_clusters(n, d, seed, noise=0.02):
    """Unit vectors in 8 very tight clusters: neighbour gaps far below the TF32 error bound."""
    rng = np.random.default_rng(seed)
    cent = np.random.default_rng(99).standard_normal((8, d))
    x = cent[rng.integers(0, 8, n)] + noise * rng.standard_normal((n, d))
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("tight", [False, True])
def test_screen_probe_one_direction(tight):
    """precision="auto": the leading rows of the first search of a fit are a probe; where the
    proof fails for most of them (tight clusters) the rest runs on 3xTF32 keys.  Exact either way."""
    from kiez_b200 import B200

    nq, ny, d, k = 3000, 4000, 64, 10
    if tight:
        q, y = _tight_clusters(nq, d, 1), _tight_clusters(ny, d, 2)
    else:
        q, y = _data(nq, ny, d, seed=5)
    algo = B200(n_candidates=k, fused=False)
    algo.SCREEN_PROBE_ROWS = 512
    algo.fit(q, y)
    dist, ind = algo.kneighbors(k=k)
    assert algo._screen_ok is (not tight), algo.search_stats
    probes = algo.search_stats["screen_probe_unverified"]
    if tight:
        # the probe ran twice (cap 16, then the boosted lists) and both left most rows unproven
        assert len(probes) == 2 and min(probes) > 0.25 and algo._screen_boost
        assert algo.search_stats["screen_rows"] == 512          # only the probe was screened
    else:
        assert algo.search_stats["screen_rows"] == nq
    want_d, want_i = O.knn_brute(q.astype(np.float64), y.astype(np.float64), k, "euclidean")
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"probe tight={tight}")
    # the verdict holds for the rest of the fit, a new fit probes again
    dist2, ind2 = algo.kneighbors(k=k, query=y, s_to_t=False)
    assert algo.search_stats["screen_rows"] == (512 if tight else nq + ny)
    want_d, want_i = O.knn_brute(y.astype(np.float64), q.astype(np.float64), k, "euclidean")
    O.assert_neighbors_match(dist2.cpu().numpy(), ind2.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"probe reverse tight={tight}")
    algo.fit(q, y)
    assert algo._screen_ok is None and not algo._screen_boost


def _dense_clusters(n, d, seed, clusters=3, noise=0.25):
    """bench.py's "hubby" distribution (unit-normalised Gaussian mixture, relative noise 0.25) with
    few clusters, so that a cluster is as densely populated as at the 1M-row size: the gap
    between the k-th and the 16th neighbour is around the proof's error bound E."""
    rng = np.random.default_rng(seed)
    cent = np.random.default_rng(77).standard_normal((clusters, d))
    x = noise * rng.standard_normal((n, d)) + cent[rng.integers(0, clusters, n)]
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("fused", [False, True])
def test_screen_probe_boosts_the_list_length(fused):
    """Neighbour gaps around E: the probe asks for longer candidate lists (`_screen_boost`)
    instead of giving up the screen; results stay exact whatever the verdicts are."""
    from kiez_b200 import B200

    nq, ny, d, k = 12000, 10000, 256, 10
    q, y = _dense_clusters(nq, d, 1), _dense_clusters(ny, d, 2)
    algo = B200(n_candidates=k, fused=fused)
    algo.SCREEN_PROBE_ROWS = 2048
    algo.FUSED_SEGMENT_MIN_ROWS = 1024
    algo.fit(q, y)
    rd, ri = algo.kneighbors(k=k, query=y, s_to_t=False)       # kiez's reverse pass first
    fd, fi = algo.kneighbors(k=k)
    probes = algo.search_stats["screen_probe_unverified"]
    print("probe fractions", probes, "boost", algo._screen_boost, "screen_ok", algo._screen_ok,
          algo.search_stats)
    assert len(probes) == (2 if algo._screen_boost else 1)
    if algo._screen_boost:
        assert probes[0] > algo.SCREEN_BOOST_UNVERIFIED and probes[1] < probes[0]
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    want_d, want_i = O.knn_brute(q64, y64, k, "euclidean")
    O.assert_neighbors_match(fd.cpu().numpy(), fi.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="fwd")
    want_d, want_i = O.knn_brute(y64, q64, k, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")


@pytest.mark.parametrize("precision", ["tf32x3", "auto"])
def test_fused_duplicate_rows_tie_with_the_thresholds(precision):
    """Several hundred identical source rows: a column's threshold (cap-th best key of the row
    sample) ties exactly with its best keys, and emits are strictly below the threshold, so the
    column may receive fewer than k rows -- it must be searched again, never padded with -1."""
    from kiez_b200 import B200

    rng = np.random.default_rng(31)
    nq, ny, d, c = 4000, 900, 64, 10
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q[::5] = q[0]                                   # 800 copies of one row, spread over the sample
    y = rng.standard_normal((ny, d)).astype(np.float32)
    y[:50] = q[0] + 0.01 * rng.standard_normal((50, d)).astype(np.float32)   # their neighbours
    algo = B200(n_candidates=c, fused=True, precision=precision)
    algo.FUSED_SEGMENT_MIN_ROWS = 512
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    (fd, fi), (rd, ri) = algo.search_both(qp, yp, c, c)
    assert int(ri.min()) >= 0 and int(fi.min()) >= 0 and bool(torch.isfinite(rd).all())
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    want_d, want_i = O.knn_brute(y64, q64, c, "euclidean")
    np.testing.assert_allclose(rd.cpu().numpy(), want_d, rtol=RTOL, atol=ATOL)
    want_d, want_i = O.knn_brute(q64, y64, c, "euclidean")
    O.assert_neighbors_match(fd.cpu().numpy(), fi.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="fwd")


@pytest.mark.parametrize("tight", [False, True])
def test_screen_probe_dual_direction(tight):
    """Dual-direction pass: the first row segment is the probe; on failure the pass starts over
    with the 3xTF32 kernels."""
    from kiez_b200 import B200

    nq, ny, d, c = 5000, 1500, 64, 10
    if tight:
        q, y = _tight_clusters(nq, d, 3), _tight_clusters(ny, d, 4)
    else:
        q, y = _data(nq, ny, d, seed=6)
    algo = B200(n_candidates=c, fused=True)
    algo.FUSED_SEGMENT_MIN_ROWS = 256
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    (fd, fi), (rd, ri) = algo.search_both(qp, yp, c, c)
    assert algo._screen_ok is (not tight), algo.search_stats
    assert algo.search_stats["screen_rows"] == (0 if tight else nq + ny)
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    want_d, want_i = O.knn_brute(q64, y64, c, "euclidean")
    O.assert_neighbors_match(fd.cpu().numpy(), fi.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="fwd")
    want_d, want_i = O.knn_brute(y64, q64, c, "euclidean")
    O.assert_neighbors_match(rd.cpu().numpy(), ri.cpu().numpy(), want_d, want_i, RTOL, ATOL, what="rev")
