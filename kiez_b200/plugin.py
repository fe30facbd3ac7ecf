"""Registers the B200 backend and the device hubness classes with an installed
``kiez`` so that the reference's own facade resolves them:

    import kiez_b200.plugin as plugin
    plugin.register()
    from kiez import Kiez
    Kiez(algorithm="B200", hubness="CSLS").fit(source, target).kneighbors(10)

kiez builds its resolvers from ``NNAlgorithm.__subclasses__()`` at import time
(kiez/neighbors/__init__.py:20-26), before an external subclass can exist, so the
classes are created here against kiez's own base classes and registered explicitly.
"""
from __future__ import annotations

from . import hubness_reduction as _hr
from .neighbors import B200Mixin


def _register(resolver, cls):
    if hasattr(resolver, "register"):
        try:
            resolver.register(cls)
            return
        except Exception:  # e.g. name conflict on re-registration
            pass
    key = resolver.normalize_cls(cls) if hasattr(resolver, "normalize_cls") else cls.__name__.lower()
    resolver.lookup_dict[key] = cls


def register():
    """Returns {"B200": cls, ...} of the classes registered with kiez."""
    import kiez  # noqa: F401  (ImportError if kiez is not installed: nothing to register with)
    import torch
    import numpy as np
    from kiez.hubness_reduction import hubness_reduction_resolver
    from kiez.neighbors import NNAlgorithm, nn_algorithm_resolver

    B200 = type("B200", (B200Mixin, NNAlgorithm), {
        "__doc__": "Exact kNN on NVIDIA B200 (kiez_b200).",
        "_ALLOWED_INPUT_TYPES": (np.ndarray, torch.Tensor),
        "__module__": __name__,
    })
    _register(nn_algorithm_resolver, B200)
    out = {"B200": B200}
    # Device hubness classes, registered under *new* names ("B200CSLS", ...) so that
    # kiez's own numpy/torch implementations stay reachable under the original ones.
    from kiez.hubness_reduction.base import HubnessReduction

    for name in ("CSLS", "LocalScaling", "MutualProximity", "DisSimLocal"):
        cls = type("B200" + name, (getattr(_hr, name), HubnessReduction), {"__module__": __name__})
        _register(hubness_reduction_resolver, cls)
        out["B200" + name] = cls
    return out
