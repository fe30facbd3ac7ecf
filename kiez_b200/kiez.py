"""``Kiez`` facade: same constructor, ``fit`` / ``kneighbors`` / ``from_path`` /
``show_*_options`` as kiez/kiez.py:14-223, resolving to the B200 backend and the
device hubness-reduction classes of this package."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Union

import numpy as np

from .hubness_reduction import (CSLS, DisSimLocal, HubnessReduction, LocalScaling,
                                MutualProximity, NoHubnessReduction)
from .neighbors import B200, NNAlgorithm
from .resolver import Resolver

nn_algorithm_resolver = Resolver(NNAlgorithm, [B200], default=B200)
hubness_reduction_resolver = Resolver(
    HubnessReduction, [NoHubnessReduction, CSLS, LocalScaling, MutualProximity, DisSimLocal],
    default=NoHubnessReduction)


def available_nn_algorithms(as_string: bool = False):
    """kiez/neighbors/util.py:18-39: try-instantiate, ImportError = unavailable."""
    out = []
    for name in sorted(nn_algorithm_resolver.options):
        try:
            nn_algorithm_resolver.make(name)
        except ImportError:
            continue
        out.append(name if as_string else nn_algorithm_resolver.lookup(name))
    return out


class Kiez:
    """Hubness-reduced nearest-neighbour search for entity alignment on B200.

    >>> k_inst = Kiez(n_candidates=10, algorithm="B200", hubness="CSLS")
    >>> k_inst.fit(source, target)
    >>> nn_dist, nn_ind = k_inst.kneighbors(5)
    """

    def __init__(self, n_candidates: int = 10, algorithm=None,
                 algorithm_kwargs: Optional[Dict[str, Any]] = None, hubness=None,
                 hubness_kwargs: Optional[Dict[str, Any]] = None):
        if not np.issubdtype(type(n_candidates), np.integer):
            raise TypeError(
                f"n_neighbors does not take {type(n_candidates)} value, enter integer value")
        if n_candidates <= 0:
            raise ValueError(f"Expected n_candidates > 0. Got {n_candidates}")
        if algorithm_kwargs is None:
            algorithm_kwargs = {"n_candidates": n_candidates}
        elif "n_candidates" not in algorithm_kwargs:
            algorithm_kwargs["n_candidates"] = n_candidates
        algorithm = nn_algorithm_resolver.make(algorithm, algorithm_kwargs)
        assert algorithm
        if hubness_kwargs is None:
            hubness_kwargs = {}
        hubness_kwargs["nn_algo"] = algorithm
        self.hubness = hubness_reduction_resolver.make(hubness, hubness_kwargs)

    @staticmethod
    def show_algorithm_options() -> List[str]:
        return available_nn_algorithms(as_string=True)

    @staticmethod
    def show_hubness_options() -> List[str]:
        return list(hubness_reduction_resolver.options)

    @property
    def algorithm(self):
        return self.hubness.nn_algo

    @algorithm.setter
    def algorithm(self, value):
        self.hubness.nn_algo = value

    def __repr__(self):
        return (f"Kiez(algorithm: {self.algorithm}, hubness: {self.hubness})"
                f" {self.algorithm._describe_source_target_fitted()}")

    @classmethod
    def from_path(cls, path: Union[str, Path]) -> "Kiez":
        with open(path) as file:
            return cls(**json.load(file))

    def fit(self, source, target=None) -> "Kiez":
        self.hubness.fit(source, target)
        return self

    def kneighbors(self, k: Optional[int] = None, return_distance: bool = True):
        dist, ind = self.hubness.kneighbors(k)
        return (dist, ind) if return_distance else ind
