"""not gpu: host-side mirror of the reference's plugin interface -- resolver names, argument
validation, k clamping, error types (tests/test_kiez.py, tests/neighbors/test_neighbor_base.py,
tests/hubness_reduction/test_wrong_inputs.py of the reference)."""
import warnings

import numpy as np
import pytest

from kiez_b200 import (CSLS, B200, DisSimLocal, Kiez, LocalScaling, MutualProximity,
                       NNAlgorithm, NoHubnessReduction)
from kiez_b200.distributed import shard_bounds
from kiez_b200.kiez import hubness_reduction_resolver, nn_algorithm_resolver
from kiez_b200.neighbors import NotFittedError, candidate_capacity
from oracle import kiez_oracle as O


class OracleNN(NNAlgorithm):
    """CPU stand-in backend (oracle arithmetic) to drive the NNAlgorithm base class."""

    valid_metrics = ("euclidean",)

    def __init__(self, n_candidates=5, metric="euclidean", p=2):
        super().__init__(n_candidates=n_candidates, metric=metric, n_jobs=None)
        self.p = p

    def _fit(self, data, is_source):
        return np.asarray(data)

    def _kneighbors(self, k, query, index, return_distance, is_self_querying):
        d, i = O.knn_brute(query, index, k, exclude_self=is_self_querying)
        return (d, i) if return_distance else i


def test_resolver_names():
    assert set(Kiez.show_hubness_options()) == {"mutualproximity", "dissimlocal", "localscaling",
                                                "no", "csls"}        # tests/test_kiez.py:145-148
    assert nn_algorithm_resolver.lookup("B200") is B200
    assert nn_algorithm_resolver.lookup("b200") is B200
    assert hubness_reduction_resolver.lookup("NoHubnessReduction") is NoHubnessReduction
    assert hubness_reduction_resolver.lookup(None) is NoHubnessReduction
    assert hubness_reduction_resolver.lookup("CSLS") is CSLS
    assert hubness_reduction_resolver.lookup(LocalScaling) is LocalScaling
    with pytest.raises(KeyError):
        hubness_reduction_resolver.lookup("nope")
    inst = OracleNN()
    assert nn_algorithm_resolver.make(inst, {"n_candidates": 3}) is inst      # instance as-is


def test_b200_needs_cuda_and_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(ImportError, match="no CPU fallback"):
        B200()
    with pytest.raises(ImportError):
        Kiez(algorithm="B200")
    assert Kiez.show_algorithm_options() == []


def test_kiez_argument_validation():
    with pytest.raises(ValueError, match="Expected"):
        Kiez(n_candidates=-1, algorithm=OracleNN())                  # tests/test_kiez.py:89-91
    with pytest.raises(TypeError, match="does not"):
        Kiez(n_candidates="1", algorithm=OracleNN())                 # :94-96
    for hub, kw in [(None, {}), ("CSLS", {}), ("MutualProximity", {"method": "empiric"}),
                    ("LocalScaling", {"method": "nicdm"}), ("DisSimLocal", {})]:
        with pytest.raises(ValueError, match="Cannot"):              # :80-86
            Kiez(algorithm=OracleNN(n_candidates=1), hubness=hub, hubness_kwargs=dict(kw))
    with pytest.raises(ValueError, match="only supports"):
        Kiez(algorithm=OracleNN(p=1, metric="minkowski"), hubness="DisSimLocal")   # :99-101
    with pytest.raises(ValueError, match="only supports"):
        Kiez(algorithm=OracleNN(metric="cosine"), hubness="DisSimLocal")           # :104-106
    assert Kiez(algorithm=OracleNN(metric="sqeuclidean"), hubness="DisSimLocal").hubness.squared
    assert not Kiez(algorithm=OracleNN(), hubness="DisSimLocal",
                    hubness_kwargs={"squared": True}).hubness.squared       # dis_sim.py:47-54
    with pytest.raises(ValueError, match="not recognized"):
        MutualProximity(method="wrong", nn_algo=OracleNN())          # test_wrong_inputs.py
    with pytest.raises(ValueError, match="Invalid method"):
        LocalScaling(method="wrong", nn_algo=OracleNN())
    assert LocalScaling(method="NICDM", nn_algo=OracleNN()).method == "nicdm"
    assert MutualProximity(method="gaussi", nn_algo=OracleNN()).method == "normal"
    assert MutualProximity(method="exact", nn_algo=OracleNN()).method == "empiric"


def test_nn_base_fit_and_k_checks():
    rng = np.random.RandomState(42)
    source, target = rng.rand(20, 5), rng.rand(50, 5)
    algo = OracleNN(n_candidates=5)
    assert "unfitted" in algo._describe_source_target_fitted()
    with pytest.raises(NotFittedError):
        algo.kneighbors()
    with pytest.raises(ValueError, match="Not implemented for input type"):
        algo.fit([[1.0]], target)
    with pytest.raises(ValueError, match="same number of features"):
        algo.fit(source, rng.rand(10, 4))
    algo.fit(source, target)
    assert "source.shape=(20, 5)" in algo._describe_source_target_fitted()
    with pytest.raises(TypeError, match="does not take"):            # test_neighbor_base.py:21-30
        algo._check_k_value(k="test", needed_space=2)
    with pytest.raises(ValueError, match="Expected"):
        algo._check_k_value(k=0, needed_space=2)
    with pytest.warns(UserWarning, match="larger than number of samples"):
        assert algo._check_k_value(k=3, needed_space=2) == 2
    d, i = algo.kneighbors()
    assert i.shape == (20, 5)
    d, i = algo.kneighbors(k=3, query=target, s_to_t=False)
    assert i.shape == (50, 3) and i.max() < 20
    # only_fit_target builds no source index (tests/test_kiez.py:22-28)
    algo2 = OracleNN()
    algo2.fit(source, target, only_fit_target=True)
    assert not hasattr(algo2, "source_index")
    # single-source mode: self is excluded only when query is None
    algo3 = OracleNN(n_candidates=3)
    algo3.fit(source)
    _, i = algo3.kneighbors()
    assert not (i == np.arange(20)[:, None]).any()
    _, i = algo3.kneighbors(query=source, s_to_t=False)
    assert (i[:, 0] == np.arange(20)).all()


def test_set_k_if_needed_warnings():
    hub = CSLS(nn_algo=OracleNN(n_candidates=5))
    with pytest.warns(UserWarning, match="No k supplied"):
        assert hub._set_k_if_needed(None) == 5
    with pytest.warns(UserWarning, match="k > n_candidates"):
        assert hub._set_k_if_needed(20) == 5
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert hub._set_k_if_needed(3) == 3
    assert "CSLS" in repr(hub) and "DisSimLocal(squared = False)" == repr(
        DisSimLocal(nn_algo=OracleNN()))


def test_candidate_capacity_and_shard_bounds():
    for c in range(1, 129):
        cap = candidate_capacity(c)
        assert cap <= 128 and cap % 8 == 0 and cap >= min(c, 128)
        if c <= 110:
            assert cap >= c + 6
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _SegmentKnobs:
    """The class-level knobs `_fused_segments` / `_fused_sample_rows` read (no device needed)."""

    FUSED_SAMPLE_DIV = 32.0
    FUSED_SEGMENT_GROWTH = 3.0
    FUSED_SEGMENT_MIN_ROWS = 16384
    FUSED_COL_CAP = 512


@pytest.mark.parametrize(("n_rows", "growth", "min_rows"), [
    (1_000_000, 3.0, 16384), (1_000_000, 2.0, 16384), (1_000_000, 1.5, 16384), (5000, 2.0, 256),
    (300, 2.0, 16384), (257, 2.0, 256), (1, 3.0, 256), (999_999, 4.0, 1000), (1_000_000, 1.0, 16384)])
def test_fused_row_segments_cover_the_rows(n_rows, growth, min_rows):
    """Row segments of the dual-direction pass (B200._fused_segments): contiguous, cover every
    row once, start on tile-pair boundaries, grow geometrically, no short tail."""
    from kiez_b200.neighbors import B200Mixin

    knobs = _SegmentKnobs()
    knobs.FUSED_SEGMENT_GROWTH, knobs.FUSED_SEGMENT_MIN_ROWS = growth, min_rows
    n_s = B200Mixin._fused_sample_rows(knobs, n_rows, 16)
    assert 1 <= n_s <= n_rows
    assert n_s == n_rows or n_s >= 8 * 16
    bounds = B200Mixin._fused_segments(knobs, n_rows, n_s)
    assert bounds[0] == 0 and bounds[-1] == n_rows
    sizes = np.diff(bounds)
    assert (sizes > 0).all()
    assert all(b % 256 == 0 for b in bounds[:-1])
    if growth <= 1.0:
        assert bounds == [0, n_rows]
    else:
        # each segment is about (growth - 1) x everything seen before it (sample included)
        seen = n_s
        for lo, hi in zip(bounds[:-2], bounds[1:-1]):
            want = max(int((growth - 1.0) * seen), min_rows, 256)
            assert want <= hi - lo < want + 256
            seen += hi - lo
        # expected emits per column ~ cap * sum(segment / seen-before): logarithmic in n / n_s
        seen, emits = n_s, 0.0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            emits += (hi - lo) / seen
            seen += hi - lo
        assert emits <= (growth - 1.0) * (np.log(max(n_rows / n_s, 1.0)) / np.log(growth) + 2.5) + 1e-9 \
            or n_rows <= min_rows * 4


def test_fused_col_cap_follows_the_list_length():
    from kiez_b200.neighbors import B200Mixin

    knobs = _SegmentKnobs()
    assert B200Mixin._fused_col_cap(knobs, 16) == 512
    knobs.FUSED_COL_CAP = 16
    assert B200Mixin._fused_col_cap(knobs, 56) == 56


@pytest.mark.parametrize(("mean", "std"), [(10.0, 3.0), (10.0, 12.5), (1.0, 4.0), (0.3, 0.9),
                                           (50.0, 7.0), (2.0, 2.0), (5.0, 100.0)])
def test_truncnorm_third_moment_matches_scipy(mean, std):
    """The closed form behind `k_skewness_truncnorm` equals what the reference computes with
    scipy (estimation.py:52-58: truncnorm(a, b).moment(3), b at the int64 maximum)."""
    from scipy import stats

    from kiez_b200.analysis import _truncnorm_third_moment

    a = (0 - mean) / std
    b = (np.iinfo(np.int64).max - mean) / std
    want = stats.truncnorm(a, b).moment(3)
    assert _truncnorm_third_moment(mean, std) == pytest.approx(want, rel=1e-9, abs=1e-12)
    assert np.isnan(_truncnorm_third_moment(3.0, 0.0))
