"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64) restatement of the kiez hot path.

This file is the parity *checker* for the CUDA path.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``; the product package ``kiez_b200``
never imports it and has no CPU fallback.

PARITY PINNING.  The reference holds known-answer vectors only for
``kiez.analysis.hubness_score`` (tests/analysis/test_estimation.py:38-98,
tests/nn_ind.npy + tests/expected_k{2,5,10,50}_hub_scores.pkl).  For kNN and
the rescalers no reference test pins values, so this restatement is pinned
against *outputs of the reference itself*, generated in the authoring container
by ``oracle/make_golden.py`` (which imports /root/reference under the shims in
``oracle/ref_shim.py``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against those
fixtures; ``tests/test_oracle_vs_reference.py`` re-checks live when
/root/reference is mounted.

Where the arithmetic lives in a third-party dependency: the exact kNN is
scikit-learn (reference pins 1.3.2, poetry.lock:1258-1259; this image has
1.9.0) called at kiez/neighbors/exact/sklearn_nearest_neighbors.py:83-101.
``knn_brute`` restates its published brute-force algorithm
(sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp: squared
euclidean via ||x||^2 - 2 x.y + ||y||^2 clamped at 0, sqrt at the end;
sklearn/metrics/pairwise.py cosine_distances: 1 - normalize(X) normalize(Y)^T
clipped to [0,2]); ``knn_sklearn`` performs the very same library call the
reference performs, and the two are cross-checked in the tests.

All functions take/return numpy arrays; distances are float64, indices int64,
exactly as the reference's SklearnNN path returns them.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

# ---------------------------------------------------------------------------
# exact kNN  (kiez/neighbors/neighbor_algorithm_base.py:53-136 +
#             kiez/neighbors/exact/sklearn_nearest_neighbors.py:83-101)
# ---------------------------------------------------------------------------

VALID_METRICS = ("euclidean", "minkowski", "l2", "sqeuclidean", "cosine")


def _pairwise(query: np.ndarray, index: np.ndarray, metric: str) -> np.ndarray:
    """Full (nq, ny) float64 distance matrix, the way sklearn's brute path does.

    euclidean/sqeuclidean: sklearn `_argkmin.pyx.tp:471-510` (expanded form,
    clamp at 0) and `:285-296` (sqrt for euclidean).
    cosine: sklearn `metrics/pairwise.py:1136-1182` (cosine_distances).
    """
    q = np.asarray(query, dtype=np.float64)
    y = np.asarray(index, dtype=np.float64)
    if metric in ("euclidean", "minkowski", "l2", "sqeuclidean"):
        qq = np.einsum("ij,ij->i", q, q)[:, None]
        yy = np.einsum("ij,ij->i", y, y)[None, :]
        d2 = qq - 2.0 * (q @ y.T) + yy
        np.maximum(d2, 0.0, out=d2)
        return d2 if metric == "sqeuclidean" else np.sqrt(d2)
    if metric == "cosine":
        qn = np.sqrt(np.einsum("ij,ij->i", q, q))
        yn = np.sqrt(np.einsum("ij,ij->i", y, y))
        qn[qn == 0.0] = 1.0  # sklearn.preprocessing.normalize leaves zero rows
        yn[yn == 0.0] = 1.0
        s = (q / qn[:, None]) @ (y / yn[:, None]).T
        d = 1.0 - s
        np.clip(d, 0.0, 2.0, out=d)
        return d
    raise ValueError(f"metric {metric!r} not supported by the oracle")


def knn_brute(query, index, k, metric="euclidean", exclude_self=False,
              chunk_rows=2048):
    """k nearest rows of ``index`` for each row of ``query`` (ascending).

    ``exclude_self``: sklearn's ``kneighbors(X=None)`` semantics
    (sklearn/neighbors/_base.py:821-826,937-958): search k+1 and drop the row's
    own index.  Ties are broken by lower index (stable argsort) -- the
    reference leaves tie order unspecified (heap / argpartition).
    """
    query = np.asarray(query)
    index = np.asarray(index)
    nq = query.shape[0]
    dist = np.empty((nq, k), dtype=np.float64)
    ind = np.empty((nq, k), dtype=np.int64)
    for lo in range(0, nq, chunk_rows):
        hi = min(nq, lo + chunk_rows)
        d = _pairwise(query[lo:hi], index, metric)
        if exclude_self:
            d[np.arange(hi - lo), np.arange(lo, hi)] = np.inf
        order = np.argsort(d, axis=1, kind="stable")[:, :k]
        ind[lo:hi] = order
        dist[lo:hi] = np.take_along_axis(d, order, axis=1)
    return dist, ind


def knn_sklearn(query, index, k, metric="euclidean", exclude_self=False, n_jobs=None):
    """The library call the reference's SklearnNN makes
    (sklearn_nearest_neighbors.py:83-101), algorithm='brute'."""
    from sklearn.neighbors import NearestNeighbors

    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric=metric, n_jobs=n_jobs)
    nn.fit(index)
    if exclude_self:
        return nn.kneighbors(X=None, n_neighbors=k, return_distance=True)
    return nn.kneighbors(X=query, n_neighbors=k, return_distance=True)


# ---------------------------------------------------------------------------
# hubness reduction  (kiez/hubness_reduction/*.py, numpy branches)
# ---------------------------------------------------------------------------

def sort_topk(dist, ind, k):
    """HubnessReduction._sort, numpy branch (hubness_reduction/base.py:80-87)."""
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(dist, order, axis=1), np.take_along_axis(ind, order, axis=1)


def csls_transform(fwd_dist, fwd_ind, rev_dist):
    """CSLS.transform (csls.py:85-96)."""
    r_train = rev_dist.mean(axis=1)
    r_test = fwd_dist.mean(axis=1).reshape(-1, 1)
    return 2 * fwd_dist - r_test - r_train[fwd_ind]


def local_scaling_transform(fwd_dist, fwd_ind, rev_dist, method="standard"):
    """LocalScaling.transform (local_scaling.py:129-151)."""
    method = method.lower()
    if method in ("ls", "standard"):
        r_t = rev_dist[:, -1]
        r_s = fwd_dist[:, -1].reshape(-1, 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            return 1.0 - np.exp(-1 * fwd_dist ** 2 / (r_s * r_t[fwd_ind]))
    if method == "nicdm":
        r_t = rev_dist.mean(axis=1)
        r_s = fwd_dist.mean(axis=1).reshape(-1, 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            return fwd_dist / np.sqrt(r_s * r_t[fwd_ind])
    raise ValueError(f"Internal: Invalid method {method}. Try 'ls' or 'nicdm'.")


def _norm_sf(x, mu, sd):
    """scipy.stats.norm.sf(x, mu, sd) = 0.5 erfc((x-mu)/(sd sqrt2)); this is the
    *numpy* branch of the reference (mutual_proximity.py:177-182), ddof=0."""
    from scipy.special import erfc

    with np.errstate(divide="ignore", invalid="ignore"):
        z = (x - mu) / sd
    return 0.5 * erfc(z / math.sqrt(2.0))


def mp_gaussian_transform(fwd_dist, fwd_ind, rev_dist):
    """MutualProximity(method='normal'): _fit (mutual_proximity.py:101-103) +
    transform (:166-183), numpy branch: nanmean / nanstd(ddof=0) / norm.sf."""
    mu_t = np.nanmean(rev_dist, axis=1)
    sd_t = np.nanstd(rev_dist, axis=1)
    mu = np.nanmean(fwd_dist, axis=1).reshape(-1, 1)
    sd = np.nanstd(fwd_dist, axis=1).reshape(-1, 1)
    p1 = _norm_sf(fwd_dist, mu, sd)
    p2 = _norm_sf(fwd_dist, mu_t[fwd_ind], sd_t[fwd_ind])
    return 1 - p1 * p2


def mp_empiric_transform(fwd_dist, fwd_ind, rev_dist, rev_ind):
    """MutualProximity(method='empiric') transform (mutual_proximity.py:185-212).

    The reference builds, per (i, j), an O(max_ind) table indexed by *source*
    ids (rev_ind) and reads it at the query's candidate *target* ids
    (fwd_ind[i]) -- an index-space mix inherited from scikit-hubness that is
    reproduced here as-is, without the O(max_ind) table: d_j[j, l] is
    rev_dist[c_j, p] if fwd_ind[i, l] == rev_ind[c_j, p] (last p wins), else
    rev_dist[c_j, -1] + 1e-6.
    """
    n, c = fwd_dist.shape
    out = np.empty_like(fwd_dist)
    for i in range(n):
        d_i = fwd_dist[i]
        cand = fwd_ind[i]
        rd = rev_dist[cand]                     # (c, c_rev)
        ri = rev_ind[cand]                      # (c, c_rev)
        d_j = np.repeat((rd[:, -1] + 1e-6)[:, None], c, axis=1)   # (c, c)
        match = ri[:, :, None] == cand[None, None, :]             # (c, c_rev, c)
        for p in range(ri.shape[1]):            # later p overwrites earlier
            m = match[:, p, :]
            d_j = np.where(m, rd[:, p][:, None], d_j)
        d = d_i[:, None]
        out[i] = 1.0 - np.sum((d_i[None, :] > d) & (d_j > d), axis=1) / c
    return out


def dsl_fit(rev_ind, source, target):
    """DisSimLocal._fit (dis_sim.py:95-108)."""
    source = np.asarray(source, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    centroids = source[rev_ind].mean(axis=1)
    x = target - centroids
    return centroids, np.einsum("ij,ij->i", x, x)


def dsl_transform(fwd_ind, query, target, target_dist_to_centroids, squared):
    """DisSimLocal.transform (dis_sim.py:139-181); ignores neigh_dist; sklearn
    euclidean_distances(squared=True) = expanded form clamped at 0."""
    query = np.asarray(query, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    tn = target[fwd_ind]                                          # (n, c, d)
    qq = np.einsum("ij,ij->i", query, query)[:, None]
    tt = np.einsum("ijk,ijk->ij", tn, tn)
    out = qq - 2.0 * np.einsum("ik,ijk->ij", query, tn) + tt
    np.maximum(out, 0.0, out=out)
    centroids = tn.mean(axis=1)
    smc = query - centroids
    out -= np.einsum("ij,ij->i", smc, smc).reshape(-1, 1)
    out -= target_dist_to_centroids[fwd_ind]
    mn = out.min()
    if mn < 0.0:
        out += -mn
    if not squared:
        out **= 1 / 2
    return out


def dsl_squared_flag(metric: str, squared: bool = True, p: int = 2) -> bool:
    """DisSimLocal.__init__ (dis_sim.py:44-61)."""
    if metric in ("euclidean", "minkowski"):
        if p != 2:
            raise ValueError("DisSimLocal only supports squared Euclidean distances.")
        return False
    if metric == "sqeuclidean":
        return True
    raise ValueError(
        f"DisSimLocal only supports squared Euclidean distances, not metric={metric}."
    )


HUBNESS = ("no", "csls", "ls", "nicdm", "mp_gaussian", "mp_empiric", "dsl")


def kiez_kneighbors(source, target=None, *, metric="euclidean", hubness="csls",
                    n_candidates=10, k=None, knn=knn_brute, squared=True):
    """End-to-end restatement of Kiez.fit + Kiez.kneighbors
    (kiez/kiez.py:160-223 -> hubness_reduction/base.py:33-105).

    Returns (dist (n,k) float64, ind (n,k) int64).
    """
    single = target is None
    if single:
        target = source
    c = n_candidates
    k = c if k is None else min(k, c)
    if hubness in (None, "no"):
        # NoHubnessReduction.fit/kneighbors (base.py:114-122)
        kk = min(k, target.shape[0])
        return knn(source, target, kk, metric, exclude_self=single)
    if c == 1:
        raise ValueError("Cannot perform hubness reduction with a single candidate per query!")
    # reverse pass (base.py:37-42): query=target passed explicitly => self NOT excluded
    c_rev = min(c, source.shape[0])
    rev_dist, rev_ind = knn(target, source, c_rev, metric, exclude_self=False)
    # forward pass (base.py:92-94): query=None => self excluded in single-source mode
    c_fwd = min(c, target.shape[0])
    fwd_dist, fwd_ind = knn(source, target, c_fwd, metric, exclude_self=single)
    if hubness == "csls":
        out = csls_transform(fwd_dist, fwd_ind, rev_dist)
    elif hubness in ("ls", "standard"):
        out = local_scaling_transform(fwd_dist, fwd_ind, rev_dist, "ls")
    elif hubness == "nicdm":
        out = local_scaling_transform(fwd_dist, fwd_ind, rev_dist, "nicdm")
    elif hubness == "mp_gaussian":
        out = mp_gaussian_transform(fwd_dist, fwd_ind, rev_dist)
    elif hubness == "mp_empiric":
        out = mp_empiric_transform(fwd_dist, fwd_ind, rev_dist, rev_ind)
    elif hubness == "dsl":
        sq = dsl_squared_flag(metric, squared)
        _, d2c = dsl_fit(rev_ind, source, target)
        out = dsl_transform(fwd_ind, source, target, d2c, sq)
    else:
        raise ValueError(f"unknown hubness {hubness!r}")
    return sort_topk(out, fwd_ind, k)


# ---------------------------------------------------------------------------
# analysis  (kiez/analysis/estimation.py)
# ---------------------------------------------------------------------------

def k_occurrence(nn_ind, k=None):
    """estimation.py:272-295: bincount of the first k columns, negatives
    dropped, minlength = number of *query rows* (named n_train there)."""
    nn_ind = np.asarray(nn_ind)
    kn = nn_ind
    if k is None:
        k = nn_ind.shape[1]
    elif k < kn.shape[1]:
        kn = kn[:, :k]
    elif k > kn.shape[1]:
        k = nn_ind.shape[1]
        warnings.warn(f"k > nn_ind.shape[1], k will be set to {k}", stacklevel=2)
    mask = kn < 0
    if np.any(mask):
        kn = kn[~mask]
    return np.bincount(kn.astype(int).ravel(), minlength=nn_ind.shape[0]), k


def truncnorm_third_moment(mean, std):
    """estimation.py:37-58: stats.truncnorm(a, b).moment(3) with a=(0-mean)/std,
    b=(int64max-mean)/std.  Closed form (scipy truncnorm._munp recurrence
    m_k = pdf(a) a^(k-1) - pdf(b) b^(k-1) + (k-1) m_(k-2), m_-1=0, m_0=1);
    pdf(b) underflows to 0 for the int64max bound."""
    a = (0 - mean) / std
    # log-space one-sided normalisation: pdf(a) = phi(a) / sf(a)
    sf = 0.5 * math.erfc(a / math.sqrt(2.0))
    pa = math.exp(-0.5 * a * a) / math.sqrt(2.0 * math.pi) / sf
    m1 = pa
    return pa * a * a + 2.0 * m1


def hubness_score(nn_ind, target_samples, *, k=None, hub_size=2.0,
                  return_value="all_but_gini", store_k_occurrence=False):
    """estimation.py:197-351 restated (all measures)."""
    occ, k = k_occurrence(nn_ind, k)
    n_test = target_samples
    n = occ.size
    mean = occ.mean()
    x0 = occ - mean
    m2 = np.mean(x0 ** 2)
    m3 = np.mean(x0 ** 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        skew = m3 / m2 ** 1.5                     # scipy.stats.skew, bias=True
    std1 = occ.std(ddof=1)
    skew_tn = truncnorm_third_moment(mean, std1)
    if return_value in ("gini", "all"):
        srt = np.sort(occ).astype(np.int64)
        # sum_ij |x_i - x_j| = 2 sum_i (2i - n + 1) x_(i)   (integer-exact)
        num = 2 * int(np.sum((2 * np.arange(n, dtype=np.int64) - n + 1) * srt))
        gini = num / (2 * n * int(occ.sum()))
    else:
        gini = np.nan
    robin = 0.5 * float(np.sum(np.abs(x0))) / float(np.sum(occ))
    atkinson = float(1.0 - 1.0 / mean * np.mean(occ ** 0.5) ** 2)
    antihubs = np.argwhere(occ == 0).ravel()
    hubs = np.argwhere(occ >= hub_size * k).ravel()
    res = {
        "k_skewness": skew,
        "k_skewness_truncnorm": skew_tn,
        "atkinson": atkinson,
        "gini": gini,
        "robinhood": robin,
        "antihubs": antihubs,
        "antihub_occurrence": antihubs.size / n,
        "hubs": hubs,
        "hub_occurrence": occ[hubs].sum() / k / n_test,
        "groupie_ratio": occ.max() / n_test / k,
    }
    if store_k_occurrence:
        res["k_occurrence"] = occ
    if return_value == "all":
        return res
    if return_value == "all_but_gini":
        del res["gini"]
        return res
    return res[return_value]


def hits(nn_ind, gold, k=(1, 5, 10)):
    """kiez/evaluate/eval_metrics.py:23-61 for array input: gold[i] is the true
    target id of source row i (or a dict source->target)."""
    if isinstance(gold, dict):
        rows = np.fromiter(gold.keys(), dtype=np.int64)
        want = np.fromiter(gold.values(), dtype=np.int64)
    else:
        gold = np.asarray(gold)
        rows, want = np.arange(len(gold)), gold
    res = {}
    for kk in k:
        res[kk] = float(np.mean((nn_ind[rows, :kk] == want[:, None]).any(axis=1)))
    return res


# ---------------------------------------------------------------------------
# tolerance-aware comparison used by every parity test
# ---------------------------------------------------------------------------

def _row_mismatch(dist, ind, ref_dist, ref_ind, rtol, atol):
    """None if the row matches up to tie order, else a message."""
    with np.errstate(invalid="ignore"):
        close = np.isclose(dist, ref_dist, rtol=rtol, atol=atol, equal_nan=True)
    if not close.all():
        c = int(np.flatnonzero(~close)[0])
        return f"distance mismatch at col {c}: got {dist[c]!r} want {ref_dist[c]!r}"
    for cpos in np.flatnonzero(ind != ref_ind):
        # distances agree, ids differ: legal only inside a run of tied oracle
        # distances (same id set), or at the last column, where the cut
        # between the c-th and the (c+1)-th neighbour is itself a tie
        d = ref_dist
        lo = cpos
        while lo > 0 and np.isclose(d[lo - 1], d[cpos], rtol=rtol, atol=atol):
            lo -= 1
        hi = cpos
        while hi + 1 < d.size and np.isclose(d[hi + 1], d[cpos], rtol=rtol, atol=atol):
            hi += 1
        if hi == d.size - 1:
            continue
        if hi == lo:
            return (f"index mismatch at col {cpos}: got {ind[cpos]} want {ref_ind[cpos]} "
                    f"and the oracle distance {d[cpos]!r} is not tied")
        if set(ind[lo:hi + 1].tolist()) != set(ref_ind[lo:hi + 1].tolist()):
            return (f"tied run [{lo},{hi}] holds different ids: {ind[lo:hi + 1]} vs "
                    f"{ref_ind[lo:hi + 1]}")
    return None


def count_mismatched_rows(dist, ind, ref_dist, ref_ind, rtol=1e-5, atol=1e-9):
    """Number of rows that differ from the oracle beyond the tie tolerance of
    `assert_neighbors_match` (bench.py's `parity_check.mismatch`), and the first message."""
    dist = np.asarray(dist, dtype=np.float64)
    ref_dist = np.asarray(ref_dist, dtype=np.float64)
    ind, ref_ind = np.asarray(ind), np.asarray(ref_ind)
    if dist.shape != ref_dist.shape or ind.shape != ref_ind.shape:
        return dist.shape[0], f"shape {dist.shape} vs {ref_dist.shape}"
    with np.errstate(invalid="ignore"):
        suspicious = ~np.isclose(dist, ref_dist, rtol=rtol, atol=atol, equal_nan=True)
    suspicious |= ind != ref_ind
    bad, first = 0, None
    for r in np.flatnonzero(suspicious.any(axis=1)):
        msg = _row_mismatch(dist[r], ind[r], ref_dist[r], ref_ind[r], rtol, atol)
        if msg is not None:
            bad += 1
            first = first or f"row {int(r)}: {msg}"
    return bad, first


def assert_neighbors_match(dist, ind, ref_dist, ref_ind, rtol=1e-5, atol=1e-9,
                           what="", max_bad_rows=0.0):
    """Indices must be identical except where the oracle's distances at the
    differing positions are tied within tolerance (BASELINE.json north_star:
    "Neighbour indices must match exactly, except where adjacent distances
    differ by less than the stated tolerance (1e-5 relative, fp32)");
    distances must agree to rtol everywhere.  Returns the number of index
    positions that differ (all of them inside tolerated ties).

    ``max_bad_rows`` (fraction of rows) exists for MutualProximity 'empiric'
    only: its output is a count of strict ``>`` comparisons in steps of 1/c
    (mutual_proximity.py:209-212), so a last-ulp difference between two tied
    distances flips a whole step -- the reference disagrees with *itself*
    across BLAS summation orders there.  Rows beyond the budget still fail.
    """
    dist = np.asarray(dist, dtype=np.float64)
    ref_dist = np.asarray(ref_dist, dtype=np.float64)
    ind = np.asarray(ind)
    ref_ind = np.asarray(ref_ind)
    assert dist.shape == ref_dist.shape, f"{what}: shape {dist.shape} vs {ref_dist.shape}"
    assert ind.shape == ref_ind.shape, f"{what}: shape {ind.shape} vs {ref_ind.shape}"
    with np.errstate(invalid="ignore"):
        suspicious = ~np.isclose(dist, ref_dist, rtol=rtol, atol=atol, equal_nan=True)
    suspicious |= ind != ref_ind
    bad = []
    for r in np.flatnonzero(suspicious.any(axis=1)):
        msg = _row_mismatch(dist[r], ind[r], ref_dist[r], ref_ind[r], rtol, atol)
        if msg is not None:
            bad.append((int(r), msg))
    if len(bad) > max_bad_rows * dist.shape[0]:
        r, msg = bad[0]
        raise AssertionError(
            f"{what}: {len(bad)} of {dist.shape[0]} rows differ beyond tie tolerance "
            f"(budget {max_bad_rows:.2%}); first: row {r}: {msg}")
    return int((ind != ref_ind).sum())
