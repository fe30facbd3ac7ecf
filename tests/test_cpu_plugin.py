"""not gpu: `kiez_b200.plugin.register()` against the reference's OWN facade (loaded through the
import shims of oracle/ref_shim.py; skipped where no copy of the reference is present) -- the
registration BASELINE.json's north star names: the backend "registers as a new kiez.neighbors
NNAlgorithm (e.g. algorithm="B200") behind the unchanged Kiez(...) API"; and `Kiez.from_path`
with the reference's tests/example_conf.json shape (kiez/kiez.py:131-158)."""
import json

import numpy as np
import pytest

from oracle import ref_shim

needs_reference = pytest.mark.skipif(not ref_shim.reference_available(),
                                     reason="no copy of the reference (baseline/_ref) present")


@needs_reference
def test_register_resolves_b200_through_the_reference_resolvers():
    kiez = ref_shim.load_reference()
    import torch

    from kiez.hubness_reduction import hubness_reduction_resolver
    from kiez.hubness_reduction.base import HubnessReduction
    from kiez.neighbors import NNAlgorithm, nn_algorithm_resolver

    import kiez_b200.plugin as plugin

    classes = plugin.register()
    assert set(classes) == {"B200", "B200CSLS", "B200LocalScaling", "B200MutualProximity",
                            "B200DisSimLocal"}
    b200 = classes["B200"]
    assert issubclass(b200, NNAlgorithm)            # the reference's own base class
    for name in ("B200", "b200"):
        assert nn_algorithm_resolver.lookup(name) is b200
    for name in ("CSLS", "LocalScaling", "MutualProximity", "DisSimLocal"):
        cls = hubness_reduction_resolver.lookup("B200" + name)
        assert cls is classes["B200" + name] and issubclass(cls, HubnessReduction)
        # the reference's own numpy/torch implementations stay reachable under the old names
        assert hubness_reduction_resolver.lookup(name).__module__.startswith("kiez.")
    assert plugin.register()["B200"].__name__ == "B200"          # idempotent
    if not torch.cuda.is_available():
        # kiez probes backends by construction: ImportError = unavailable (kiez/kiez.py:118-122,
        # kiez/neighbors/util.py:31-38); there is no CPU fallback to construct
        with pytest.raises(ImportError, match="no\\s+CPU fallback|CUDA"):
            kiez.Kiez(algorithm="B200", hubness="B200CSLS")
        with pytest.raises(ImportError):
            b200()


@needs_reference
def test_registered_classes_keep_the_reference_contract():
    """What kiez's facade reads from a backend class."""
    ref_shim.load_reference()
    import torch

    import kiez_b200.plugin as plugin

    b200 = plugin.register()["B200"]
    assert np.ndarray in b200._ALLOWED_INPUT_TYPES and torch.Tensor in b200._ALLOWED_INPUT_TYPES
    assert {"euclidean", "sqeuclidean", "cosine", "minkowski"} <= set(b200.valid_metrics)
    for hook in ("_fit", "_kneighbors", "fit", "kneighbors"):
        assert callable(getattr(b200, hook))
    assert not getattr(b200, "__abstractmethods__", None)


def test_from_path_reads_a_reference_style_config(tmp_path):
    """Kiez.from_path (kiez/kiez.py:154-158) with the shape of the reference's
    tests/example_conf.json: constructor kwargs as JSON.  Without a CUDA device the backend's
    constructor raises ImportError (no CPU fallback) -- after the JSON was parsed and resolved."""
    import torch

    from kiez_b200 import Kiez

    conf = {"n_candidates": 7, "algorithm": "B200",
            "algorithm_kwargs": {"metric": "cosine"},
            "hubness": "LocalScaling", "hubness_kwargs": {"method": "NICDM"}}
    path = tmp_path / "conf.json"
    path.write_text(json.dumps(conf))
    if not torch.cuda.is_available():
        with pytest.raises(ImportError):
            Kiez.from_path(path)
        bad = tmp_path / "bad.json"
        bad.write_text(json.dumps(dict(conf, algorithm="NoSuchBackend")))
        with pytest.raises((KeyError, ValueError)):
            Kiez.from_path(bad)
        return
    inst = Kiez.from_path(str(path))
    assert inst.algorithm.n_candidates == 7 and inst.algorithm.metric == "cosine"
    assert type(inst.hubness).__name__ == "LocalScaling" and inst.hubness.method == "nicdm"
