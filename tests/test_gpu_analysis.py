"""-m gpu: hubness_score / hits on device against the reference's own known-answer vectors
(tests/analysis/test_estimation.py) and the oracle."""
import os

import numpy as np
import pytest

import _golden
from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _z():
    return np.load(os.path.join(_golden.GOLDEN_DIR, "hubness_score.npz"))


@pytest.mark.parametrize("k", [2, 5, 10, 50])
def test_known_answers(k):
    from kiez_b200 import hubness_score

    z = _z()
    nn_ind = z["nn_ind"].astype(np.int64)
    res = hubness_score(nn_ind, 1000, k=k, return_value="all", store_k_occurrence=True)
    keys = [f[len(f"k{k}__"):] for f in z.files if f.startswith(f"k{k}__")]
    assert len(keys) >= 10
    for key in keys:
        want = z[f"k{k}__{key}"]
        if want.ndim:
            np.testing.assert_array_equal(res[key], want)      # integer work: bit-exact
        else:
            assert res[key] == pytest.approx(float(want), rel=1e-9), key


def test_toy_negative_and_warning():
    from kiez_b200 import hubness_score

    z = _z()
    s = hubness_score(z["toy_nn"], 5)
    np.testing.assert_almost_equal(s["k_skewness"], float(z["toy_k_skewness"]), decimal=10)
    assert hubness_score(np.array([[1, 2, 3], [-1, 4, 5]]), 5) is not None
    with pytest.warns(UserWarning, match="k > nn_ind"):
        hubness_score(np.array([[1, 2, 3], [-1, 4, 5]]), 5, k=10)
    with pytest.raises(ValueError, match="no negative"):
        hubness_score(np.array([[np.inf], [0]]), 1)


def test_large_random_vs_oracle():
    from kiez_b200 import hubness_score

    rng = np.random.default_rng(3)
    nn = rng.integers(0, 200000, (150000, 10))
    nn[rng.random(nn.shape) < 0.5] //= 97
    want = O.hubness_score(nn, 200000, k=7, return_value="all", store_k_occurrence=True)
    got = hubness_score(torch.from_numpy(nn).cuda(), 200000, k=7, return_value="all",
                        store_k_occurrence=True)
    for key, val in want.items():
        g = got[key]
        if isinstance(val, np.ndarray):
            np.testing.assert_array_equal(g.cpu().numpy(), val)
        else:
            assert g == pytest.approx(val, rel=1e-9), key
    assert int(got["k_occurrence"].sum()) == 150000 * 7       # checksum of the histogram


def test_hits():
    from kiez_b200 import hits

    nn = np.array([[3, 1, 2], [0, 2, 1], [1, 0, 2]])
    assert hits(nn, [3, 1, 2], k=(1, 2, 3)) == {1: 1 / 3, 2: 1 / 3, 3: 1.0}
    assert hits(nn, {0: 3, 2: 0}, k=(1, 2)) == {1: 0.5, 2: 1.0}
    rng = np.random.default_rng(0)
    big = rng.integers(0, 1000, (50000, 10))
    gold = rng.integers(0, 1000, 50000)
    assert hits(big, gold, k=(1, 5, 10)) == pytest.approx(O.hits(big, gold, k=(1, 5, 10)))
    # the reference's corner cases (eval_metrics.py:8-12,53-61): k=None -> [1, 5, 10]; the
    # denominator is len(gold), so a key that is not a row counts as a miss; list / dict input
    doc = np.array([[1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 5, 6]])
    assert hits(doc, {0: 2, 1: 4, 2: 3, 3: 4}) == {1: 0.5, 5: 1.0, 10: 1.0}     # its docstring example
    assert hits(doc.tolist(), {0: 2, 1: 4, 2: 3, 3: 4}, k=[5, 1]) == {1: 0.5, 5: 1.0}
    assert hits(doc, {0: 2, 1: 4, 99: 7, -1: 3}, k=[1, 3]) == {1: 0.0, 3: 0.5}
    assert hits({10: [1, 2, 3], 11: [2, 3, 4]}, {10: 2, 11: 2, 12: 0}, k=[1, 2]) == \
        {1: 1 / 3, 2: 2 / 3}
    assert hits(doc, np.array([1, 4]), k=[1, 3]) == {1: 0.5, 3: 1.0}            # short gold array
