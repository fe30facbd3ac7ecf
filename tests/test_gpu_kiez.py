"""-m gpu: Kiez(algorithm="B200", hubness=...) end to end against (a) the committed outputs of
the real reference (tests/golden, made by oracle/make_golden.py) and (b) the CPU oracle on
fresh seeded inputs.  Reads like the reference's tests/test_kiez.py + tests/neighbors/test_faiss.py."""
import numpy as np
import pytest

import _golden
from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RTOL, ATOL = 1e-5, 5e-6

HUB = {
    "no": (None, {}),
    "csls": ("CSLS", {}),
    "ls": ("LocalScaling", {"method": "standard"}),
    "nicdm": ("LocalScaling", {"method": "nicdm"}),
    "mp_gaussian": ("MutualProximity", {"method": "normal"}),
    "mp_empiric": ("MutualProximity", {"method": "empiric"}),
    "dsl": ("DisSimLocal", {}),
}


def _kiez(c, metric, hub, impl="auto", fused="auto"):
    from kiez_b200 import B200, Kiez

    name, kw = HUB[hub]
    return Kiez(n_candidates=c,
                algorithm=B200(n_candidates=c, metric=metric, impl=impl, fused=fused),
                hubness=name, hubness_kwargs=dict(kw))


@pytest.mark.parametrize(("name", "metric", "hub"), _golden.cases())
@pytest.mark.parametrize("fused", [False, True])
def test_matches_reference_golden(name, metric, hub, fused):
    source, target, c, k, ref_dist, ref_ind = _golden.get(name, metric, hub)
    inst = _kiez(c, metric, hub, fused=fused)
    inst.fit(source, target)
    dist, ind = inst.kneighbors(k)
    assert isinstance(dist, np.ndarray) and dist.dtype == np.float64 and ind.dtype == np.int64
    O.assert_neighbors_match(dist, ind, ref_dist, ref_ind, RTOL, ATOL,
                             what=f"{name}/{metric}/{hub}",
                             max_bad_rows=0.03 if hub == "mp_empiric" else 0.0)


@pytest.mark.parametrize("hub", list(HUB))
@pytest.mark.parametrize("impl", ["tc", "tc1", "simt"])
@pytest.mark.parametrize("single", [False, True])
def test_matches_oracle_fresh_inputs(hub, impl, single):
    rng = np.random.default_rng(17)
    source = rng.standard_normal((900, 48)).astype(np.float32)
    target = None if single else rng.standard_normal((1100, 48)).astype(np.float32)
    inst = _kiez(24, "euclidean", hub, impl)
    inst.fit(torch.from_numpy(source).cuda(), None if single else torch.from_numpy(target).cuda())
    dist, ind = inst.kneighbors(10)
    assert torch.is_tensor(dist) and dist.is_cuda and ind.dtype == torch.int64
    want_d, want_i = O.kiez_kneighbors(source.astype(np.float64),
                                       None if single else target.astype(np.float64),
                                       hubness=hub, n_candidates=24, k=10)
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, RTOL, ATOL,
                             what=f"{hub}/{impl}/single={single}",
                             max_bad_rows=0.03 if hub == "mp_empiric" else 0.0)


@pytest.mark.parametrize("hub", ["csls", "ls", "nicdm", "mp_gaussian", "mp_empiric", "dsl"])
def test_transform_is_unsorted_and_matches_oracle(hub):
    """HubnessReduction.transform contract (docs/source/using_your_own.rst:11-19)."""
    rng = np.random.default_rng(23)
    source = rng.standard_normal((400, 32)).astype(np.float32)
    target = rng.standard_normal((500, 32)).astype(np.float32)
    s64, t64 = source.astype(np.float64), target.astype(np.float64)
    c = 12
    inst = _kiez(c, "euclidean", hub)
    inst.fit(source, target)
    fd, fi = inst.algorithm.kneighbors(k=c)
    got, got_i = inst.hubness.transform(fd, fi, inst.algorithm.source_)
    assert torch.equal(got_i, fi)
    rd, ri = O.knn_brute(t64, s64, c)
    wd, wi = O.knn_brute(s64, t64, c)
    np.testing.assert_array_equal(fi.cpu().numpy(), wi)
    if hub == "csls":
        want = O.csls_transform(wd, wi, rd)
    elif hub in ("ls", "nicdm"):
        want = O.local_scaling_transform(wd, wi, rd, hub)
    elif hub == "mp_gaussian":
        want = O.mp_gaussian_transform(wd, wi, rd)
    elif hub == "mp_empiric":
        want = O.mp_empiric_transform(wd, wi, rd, ri)
    else:
        want = O.dsl_transform(wi, s64, t64, O.dsl_fit(ri, s64, t64)[1], squared=False)
    got = got.cpu().numpy()
    if hub == "mp_empiric":
        assert (np.abs(got - want) > 1e-9).mean() < 0.01
    else:
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


def test_api_behaviour():
    """tests/test_kiez.py:22-55,66-96 of the reference: shapes, clamping warnings, errors."""
    from kiez_b200 import B200, Kiez

    rng = np.random.RandomState(42)
    source, target = rng.rand(20, 5), rng.rand(50, 5)
    k_inst = Kiez(n_candidates=10, algorithm="B200")
    k_inst.fit(source, target)
    assert not hasattr(k_inst.algorithm, "source_index")       # no hubness: target index only
    assert "B200" in f"{k_inst}"
    assert Kiez(n_candidates=7, algorithm="B200",
                algorithm_kwargs={"metric": "cosine"}).algorithm.n_candidates == 7
    for hub, kw in [(None, {}), ("CSLS", {}), ("MutualProximity", {"method": "empiric"}),
                    ("LocalScaling", {"method": "nicdm"}), ("DisSimLocal", {"squared": True})]:
        inst = Kiez(algorithm="B200", n_candidates=5, hubness=hub, hubness_kwargs=dict(kw))
        for tgt in (target, None):
            inst.fit(source, tgt)
            with pytest.warns(UserWarning):
                dist, neigh = inst.kneighbors()
            assert neigh.shape == (20, 5) and dist.shape == (20, 5)
            assert inst.kneighbors(return_distance=False, k=5).shape == (20, 5)
            dist, neigh = inst.kneighbors(k=1)
            assert neigh.shape == (20, 1)
            with pytest.warns(UserWarning):
                dist, neigh = inst.kneighbors(k=20)
            assert neigh.shape == (20, 5)
        with pytest.raises(ValueError, match="Cannot"):
            Kiez(algorithm="B200", n_candidates=1, hubness=hub, hubness_kwargs=dict(kw))
    with pytest.raises(ValueError, match="only supports"):
        Kiez(algorithm=B200(metric="cosine"), hubness="DisSimLocal")
    with pytest.raises(ValueError, match="same number of features"):
        Kiez(algorithm="B200").fit(source, rng.rand(10, 4))
    with pytest.raises(ValueError, match="Not implemented for input type"):
        Kiez(algorithm="B200").fit([[1.0, 2.0]], target)
    assert Kiez(algorithm=B200(metric="sqeuclidean"), hubness="DisSimLocal").hubness.squared
    assert "b200" in Kiez.show_algorithm_options()


CONFIG_SHAPED = [
    # BASELINE.json configs at oracle-sized row counts (same d, c, k, hubness as the named config)
    ("c2", 3000, 3500, 256, 50, 10, "csls"),
    ("c3-mp", 2500, 3000, 256, 100, 10, "mp_gaussian"),
    ("c3-ls", 2500, 3000, 256, 100, 10, "ls"),
    ("c3-nicdm", 2500, 3000, 256, 100, 10, "nicdm"),
    ("c4", 5000, 6000, 256, 10, 10, "csls"),
    ("c5-dsl", 3000, 9000, 128, 50, 10, "dsl"),
    ("c5-nicdm", 3000, 9000, 128, 50, 10, "nicdm"),
]


CONFIG_SEEDS = {name: 101 + i for i, (name, *_rest) in enumerate(CONFIG_SHAPED)}


@pytest.mark.parametrize(("name", "n", "m", "d", "c", "k", "hub"), CONFIG_SHAPED)
def test_config_shaped_parity(name, n, m, d, c, k, hub):
    rng = np.random.default_rng(CONFIG_SEEDS[name])      # fixed: str hashes are salted per process
    source = rng.standard_normal((n, d)).astype(np.float32)
    target = rng.standard_normal((m, d)).astype(np.float32)
    want_d, want_i = O.kiez_kneighbors(source.astype(np.float64), target.astype(np.float64),
                                       hubness=hub, n_candidates=c, k=k, knn=O.knn_sklearn)
    for fused in (False, True) if c <= 50 else (False,):
        inst = _kiez(c, "euclidean", hub, fused=fused)
        inst.fit(source, target)
        dist, ind = inst.kneighbors(k)
        O.assert_neighbors_match(dist, ind, want_d, want_i, RTOL, ATOL, what=f"{name} fused={fused}")


def test_full_size_properties_c4_like():
    """Size-independent properties at a size the oracle cannot run in full: sortedness,
    id range, no duplicate ids, idempotence (same result twice), and oracle parity on a
    random row sample (forward candidates are recomputed by brute force for those rows)."""
    from kiez_b200 import hubness_score

    n = m = 60000
    d, c, k = 256, 10, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    source = torch.randn((n, d), generator=g, device="cuda")
    target = torch.randn((m, d), generator=g, device="cuda")
    inst = _kiez(c, "euclidean", "csls")
    inst.fit(source, target)
    dist, ind = inst.kneighbors(k)
    dist2, ind2 = inst.kneighbors(k)
    assert torch.equal(ind, ind2) and torch.equal(dist, dist2)
    assert (torch.diff(dist, dim=1) >= 0).all()
    assert int(ind.min()) >= 0 and int(ind.max()) < m
    srt = torch.sort(ind, dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()
    # oracle on a row sample: CSLS needs the reverse statistics of every target the sample
    # touches, so compute the reverse pass for exactly those targets
    rows = np.random.default_rng(1).choice(n, 128, replace=False)
    s64 = source.cpu().numpy().astype(np.float64)
    t64 = target.cpu().numpy().astype(np.float64)
    fwd_d, fwd_i = O.knn_brute(s64[rows], t64, c)
    touched = np.unique(fwd_i)
    rev_d, _ = O.knn_brute(t64[touched], s64, c)
    r_train = np.zeros(m)
    r_train[touched] = rev_d.mean(axis=1)
    want = 2 * fwd_d - fwd_d.mean(axis=1, keepdims=True) - r_train[fwd_i]
    want_d, want_i = O.sort_topk(want, fwd_i, k)
    O.assert_neighbors_match(dist[rows].cpu().numpy(), ind[rows].cpu().numpy(), want_d, want_i,
                             RTOL, ATOL, what="c4-like sample")
    scores = hubness_score(ind, m, k=k, store_k_occurrence=True)
    assert int(scores["k_occurrence"].sum()) == n * k          # checksum of the histogram


def test_full_size_c4_sampled_parity():
    """BASELINE.json's metric configuration at FULL size (C4: 1M x 1M, d=256, c=k=10, CSLS) through
    the default path -- 1xTF32 screen + proof, dual-direction pass in row segments: oracle parity
    on sampled rows of BOTH directions, plus the size-independent properties (sortedness, id
    range, no duplicates, histogram checksum).  The oracle side uses the scikit-learn brute-force
    call the reference makes, on float64 copies of the same values."""
    from kiez_b200 import B200, Kiez, hubness_score

    free, _total = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs 40 GB of free device memory")
    n = m = 1_000_000
    d, c, k = 256, 10, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    source = torch.randn((n, d), generator=g, device="cuda")
    g.manual_seed(1)
    target = torch.randn((m, d), generator=g, device="cuda")
    inst = Kiez(n_candidates=c, algorithm=B200(n_candidates=c), hubness="CSLS")
    inst.fit(source, target)
    algo = inst.algorithm
    assert algo._fused_forward is not None, "C4 must take the dual-direction pass by default"
    assert algo.search_stats["screen_rows"] == n + m, "C4 must take the 1xTF32 screen by default"
    assert algo.search_stats["screen_unverified"] < 0.01 * (n + m)
    dist, ind = inst.kneighbors(k)
    assert (torch.diff(dist, dim=1) >= 0).all()
    assert int(ind.min()) >= 0 and int(ind.max()) < m
    srt = torch.sort(ind, dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()
    rev_dist, rev_ind = inst.hubness.r_dist_train_, inst.hubness.r_ind_train_
    assert (torch.diff(rev_dist, dim=1) >= 0).all()
    assert int(rev_ind.min()) >= 0 and int(rev_ind.max()) < n
    # oracle on >= 4096 rows per direction: stratified over the row segments of the pass, plus
    # EVERY row / column whose proof failed (they took the 3xTF32 re-search)
    rng = np.random.default_rng(1)
    bounds = algo._fused_segments(n, algo._fused_sample_rows(n, 16))
    per_seg = -(-4096 // (len(bounds) - 1))
    rows = np.concatenate([rng.choice(np.arange(lo, hi), min(per_seg, hi - lo), replace=False)
                           for lo, hi in zip(bounds[:-1], bounds[1:])])
    researched = {key: val.cpu().numpy() for key, val in algo.researched.items()}
    assert len(researched["rows"]) + len(researched["cols"]) == algo.search_stats["screen_unverified"]
    rows = np.unique(np.concatenate([rows, researched["rows"]]))
    s64 = source.cpu().numpy().astype(np.float64)
    t64 = target.cpu().numpy().astype(np.float64)
    fwd_d, fwd_i = O.knn_sklearn(s64[rows], t64, c, n_jobs=-1)
    got_fd, got_fi = algo.kneighbors(k=c)                      # forward kNN before the rescale
    O.assert_neighbors_match(got_fd[rows].cpu().numpy(), got_fi[rows].cpu().numpy(), fwd_d, fwd_i,
                             RTOL, ATOL, what=f"c4 forward kNN ({len(rows)} rows)")
    # reverse: 4096 sampled columns + the re-searched ones + everything 64 of the rows touch
    # (their CSLS value needs the reverse statistics of all their candidates)
    csls_rows = rows[rng.choice(len(rows), 64, replace=False)]
    touched = np.unique(fwd_i[np.isin(rows, csls_rows)])
    cols = np.unique(np.concatenate([rng.choice(m, 4096, replace=False), researched["cols"], touched]))
    want_rev_d, want_rev_i = O.knn_sklearn(t64[cols], s64, c, n_jobs=-1)
    O.assert_neighbors_match(rev_dist[cols].cpu().numpy(), rev_ind[cols].cpu().numpy(),
                             want_rev_d, want_rev_i, RTOL, ATOL,
                             what=f"c4 reverse kNN (column side, {len(cols)} columns)")
    r_train = np.zeros(m)
    r_train[cols] = want_rev_d.mean(axis=1)
    sel = np.isin(rows, csls_rows)
    want = 2 * fwd_d[sel] - fwd_d[sel].mean(axis=1, keepdims=True) - r_train[fwd_i[sel]]
    want_d, want_i = O.sort_topk(want, fwd_i[sel], k)
    O.assert_neighbors_match(dist[rows[sel]].cpu().numpy(), ind[rows[sel]].cpu().numpy(), want_d,
                             want_i, RTOL, ATOL, what="c4 forward + CSLS")
    print(f"c4 full size: {len(rows)} rows + {len(cols)} columns checked, "
          f"{len(researched['rows'])} + {len(researched['cols'])} of them re-searched")
    scores = hubness_score(ind, m, k=k, store_k_occurrence=True)
    assert int(scores["k_occurrence"].sum()) == n * k          # checksum of the histogram


def test_sort_mirrors_the_input_container():
    """HubnessReduction._sort (base.py:72-87; the reference's tests/hubness_reduction/
    test_hubness_base.py): numpy in -> numpy out, torch in -> torch out on the input's device,
    same values either way, equal to the oracle's top-k."""
    from kiez_b200 import HubnessReduction

    rng = np.random.default_rng(seed=42)
    dist = rng.random((100, 10))
    ind = rng.integers(low=0, high=200, size=(100, 10))
    np_dist, np_ind = HubnessReduction._sort(dist, ind, 10)
    assert isinstance(np_dist, np.ndarray) and isinstance(np_ind, np.ndarray)
    t_dist, t_ind = HubnessReduction._sort(torch.tensor(dist), torch.tensor(ind), 10)
    assert isinstance(t_dist, torch.Tensor) and isinstance(t_ind, torch.Tensor)
    assert not t_dist.is_cuda and not t_ind.is_cuda
    c_dist, c_ind = HubnessReduction._sort(torch.tensor(dist).cuda(), torch.tensor(ind).cuda(), 4)
    assert c_dist.is_cuda and c_dist.shape == (100, 4)
    np.testing.assert_array_equal(t_dist.numpy(), np_dist)
    np.testing.assert_array_equal(t_ind.numpy(), np_ind)
    want_d, want_i = O.sort_topk(dist, ind, 10)
    np.testing.assert_array_equal(np_dist, want_d)
    np.testing.assert_array_equal(np_ind, want_i)
    np.testing.assert_array_equal(c_dist.cpu().numpy(), want_d[:, :4])
    np.testing.assert_array_equal(c_ind.cpu().numpy(), want_i[:, :4])
