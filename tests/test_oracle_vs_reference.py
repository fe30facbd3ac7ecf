"""Live re-check of the oracle against the unmodified reference, imported under
the shims of oracle/ref_shim.py.  Skipped where /root/reference is not mounted
(the GPU box); the committed fixtures in tests/golden/ cover that case."""
import numpy as np
import pytest

from oracle import kiez_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason="/root/reference not mounted")

CASES = [
    ("no", None, {}), ("csls", "CSLS", {}), ("ls", "LocalScaling", {"method": "ls"}),
    ("nicdm", "LocalScaling", {"method": "nicdm"}),
    ("mp_gaussian", "MutualProximity", {"method": "normal"}),
    ("mp_empiric", "MutualProximity", {"method": "empiric"}),
    ("dsl", "DisSimLocal", {}),
]


@pytest.mark.parametrize(("label", "hub", "kw"), CASES)
@pytest.mark.parametrize("single", [False, True])
def test_live(label, hub, kw, single):
    rng = np.random.default_rng(123)
    s = rng.standard_normal((150, 24)).astype(np.float32).astype(np.float64)
    t = None if single else rng.standard_normal((90, 24)).astype(np.float32).astype(np.float64)
    ref_d, ref_i = ref_shim.reference_kneighbors(s, t, hubness=hub, hubness_kwargs=kw,
                                                 n_candidates=12, k=7)
    d, i = O.kiez_kneighbors(s, t, hubness=label, n_candidates=12, k=7)
    O.assert_neighbors_match(d, i, ref_d, ref_i, rtol=1e-7, atol=1e-7, what=label,
                             max_bad_rows=0.02 if label == "mp_empiric" else 0.0)


def test_live_hubness_score():
    ref_shim.load_reference()
    from kiez.analysis import hubness_score

    rng = np.random.default_rng(5)
    nn = rng.integers(0, 400, (500, 10))
    nn[rng.random(nn.shape) < 0.6] //= 7          # make it skewed
    want = hubness_score(nn, 400, k=8, return_value="all", store_k_occurrence=True)
    got = O.hubness_score(nn, 400, k=8, return_value="all", store_k_occurrence=True)
    for key, val in want.items():
        if isinstance(val, np.ndarray):
            np.testing.assert_array_equal(got[key], val)
        else:
            assert got[key] == pytest.approx(val, rel=1e-9), key
