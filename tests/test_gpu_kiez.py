"""-m gpu: Kiez(algorithm="B200", hubness=...) end to end against (a) the committed outputs of
the real reference (tests/golden, made by oracle/make_golden.py) and (b) the CPU oracle on
fresh seeded inputs.  Reads like the reference's tests/test_kiez.py + tests/neighbors/test_faiss.py."""
import numpy as np
import pytest

import _golden
from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RTOL, ATOL = 1e-5, 1e-7

HUB = {
    "no": (None, {}),
    "csls": ("CSLS", {}),
    "ls": ("LocalScaling", {"method": "standard"}),
    "nicdm": ("LocalScaling", {"method": "nicdm"}),
    "mp_gaussian": ("MutualProximity", {"method": "normal"}),
    "mp_empiric": ("MutualProximity", {"method": "empiric"}),
    "dsl": ("DisSimLocal", {}),
}


def _kiez(c, metric, hub, impl="auto"):
    from kiez_b200 import B200, Kiez

    name, kw = HUB[hub]
    return Kiez(n_candidates=c, algorithm=B200(n_candidates=c, metric=metric, impl=impl),
                hubness=name, hubness_kwargs=dict(kw))


@pytest.mark.parametrize(("name", "metric", "hub"), _golden.cases())
def test_matches_reference_golden(name, metric, hub):
    source, target, c, k, ref_dist, ref_ind = _golden.get(name, metric, hub)
    inst = _kiez(c, metric, hub)
    inst.fit(source, target)
    dist, ind = inst.kneighbors(k)
    assert isinstance(dist, np.ndarray) and dist.dtype == np.float64 and ind.dtype == np.int64
    O.assert_neighbors_match(dist, ind, ref_dist, ref_ind, RTOL, ATOL,
                             what=f"{name}/{metric}/{hub}",
                             max_bad_rows=0.03 if hub == "mp_empiric" else 0.0)


@pytest.mark.parametrize("hub", list(HUB))
@pytest.mark.parametrize("impl", ["tc", "simt"])
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
