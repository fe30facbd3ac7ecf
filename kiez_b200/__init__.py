"""kiez_b200 -- B200-native exact kNN + hubness reduction behind kiez's API.

    from kiez_b200 import Kiez
    nn_dist, nn_ind = Kiez(n_candidates=10, algorithm="B200", hubness="CSLS") \\
        .fit(source, target).kneighbors(5)

Importing the package never touches the GPU; constructing ``B200`` without a CUDA
device (or without the built library) raises ImportError -- there is no CPU path.
"""
from .analysis import hubness_score
from .evaluate import hits
from .hubness_reduction import (CSLS, DisSimLocal, HubnessReduction, LocalScaling,
                                MutualProximity, NoHubnessReduction)
from .kiez import Kiez, hubness_reduction_resolver, nn_algorithm_resolver
from .neighbors import B200, NNAlgorithm

__version__ = "0.1.0"
__all__ = [
    "Kiez", "B200", "NNAlgorithm", "HubnessReduction", "NoHubnessReduction", "CSLS",
    "LocalScaling", "MutualProximity", "DisSimLocal", "hubness_score", "hits",
    "nn_algorithm_resolver", "hubness_reduction_resolver",
]
