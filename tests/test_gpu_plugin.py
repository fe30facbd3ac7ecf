"""-m gpu: the reference's OWN facade driving the registered backend -- kiez.Kiez(algorithm="B200",
hubness="B200CSLS") from the unmodified copy under baseline/_ref (loaded through
oracle/ref_shim.py; skipped when that copy is absent) must reproduce the committed outputs of the
reference's SklearnNN path (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

import _golden
from oracle import kiez_oracle as O
from oracle import ref_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(),
                                 reason="no copy of the reference (baseline/_ref) present")]
torch = pytest.importorskip("torch")

HUB = {
    "no": (None, {}),
    "csls": ("B200CSLS", {}),
    "ls": ("B200LocalScaling", {"method": "standard"}),
    "nicdm": ("B200LocalScaling", {"method": "nicdm"}),
    "mp_gaussian": ("B200MutualProximity", {"method": "normal"}),
    "dsl": ("B200DisSimLocal", {}),
}


@pytest.fixture(scope="module")
def ref_kiez():
    kiez = ref_shim.load_reference()
    import kiez_b200.plugin as plugin

    plugin.register()
    return kiez


@pytest.mark.parametrize(("name", "metric", "hub"),
                         [c for c in _golden.cases() if c[2] in HUB])
def test_reference_facade_with_registered_backend_matches_golden(ref_kiez, name, metric, hub):
    source, target, c, k, ref_dist, ref_ind = _golden.get(name, metric, hub)
    hub_name, kw = HUB[hub]
    inst = ref_kiez.Kiez(n_candidates=c, algorithm="B200", algorithm_kwargs={"metric": metric},
                         hubness=hub_name, hubness_kwargs=dict(kw))
    assert type(inst.algorithm).__mro__[2].__module__.startswith("kiez.")   # kiez's NNAlgorithm
    inst.fit(source, target)
    dist, ind = inst.kneighbors(k)
    if hub == "no":
        # the reference's NoHubnessReduction hands the backend's result through unchanged
        # (hubness_reduction/base.py:117-122): device tensors, like its Faiss-GPU backend
        assert torch.is_tensor(dist) and dist.is_cuda and ind.dtype == torch.int64
        dist, ind = dist.cpu().numpy(), ind.cpu().numpy()
    assert isinstance(dist, np.ndarray) and ind.dtype == np.int64
    O.assert_neighbors_match(dist, ind, ref_dist, ref_ind, 1e-5, 5e-6,
                             what=f"ref facade {name}/{metric}/{hub}")


def test_reference_hubness_class_on_top_of_the_registered_backend(ref_kiez):
    """The reference's own CSLS (its torch branch: the backend returns CUDA tensors,
    hubness_reduction/base.py:43-44) over B200 candidates equals the golden of its numpy branch."""
    source, target, c, k, ref_dist, ref_ind = _golden.get("gauss", "euclidean", "csls")
    inst = ref_kiez.Kiez(n_candidates=c, algorithm="B200", hubness="CSLS")
    inst.fit(torch.from_numpy(source).cuda(), torch.from_numpy(target).cuda())
    dist, ind = inst.kneighbors(k)
    assert torch.is_tensor(dist) and dist.is_cuda
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), ref_dist, ref_ind, 1e-5, 5e-6,
                             what="reference CSLS over B200")
