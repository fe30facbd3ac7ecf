"""Helpers to iterate the committed golden fixtures (tests/golden/*.npz), which
were produced from the real reference by oracle/make_golden.py."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATASETS = ("conftest", "conftest_single", "readme", "readme_single", "gauss", "hubby")
HUBNESS = ("no", "csls", "ls", "nicdm", "mp_gaussian", "mp_empiric", "dsl")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, f"kiez_{name}.npz"))


def cases():
    """Yield (dataset, metric, hubness) for every golden result."""
    out = []
    for name in DATASETS:
        with load(name) as z:
            keys = [k for k in z.files if k.endswith("__dist")]
        for key in sorted(keys):
            metric, hub, _ = key.split("__")
            out.append((name, metric, hub))
    return out


def get(name, metric, hub):
    with load(name) as z:
        source = z["source"]
        target = z["target"] if "target" in z.files else None
        c = int(z["n_candidates"])
        k = int(z["k"])
        dist = z[f"{metric}__{hub}__dist"]
        ind = z[f"{metric}__{hub}__ind"]
    return source, target, c, k, dist, ind
