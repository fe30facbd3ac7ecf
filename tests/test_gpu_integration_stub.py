"""-m gpu: the minimal reference-side binding shown in INTEGRATION.md section 2 is executed
verbatim (only the kiez import and the library path are redirected) and checked against the
oracle, so the documented stub cannot drift from include/kiez_b200.h."""
import os
import re

import numpy as np
import pytest

from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text[text.index("## 2. The minimal stub"):]
    code = re.search(r"```python\n(.*?)```", section, re.S).group(1)
    from kiez_b200 import _lib

    code = code.replace("from kiez.neighbors.neighbor_algorithm_base import NNAlgorithm",
                        "from kiez_b200.neighbors import NNAlgorithm")
    code = code.replace('C.CDLL("libkiez_b200.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    assert "kiez_b200.neighbors import NNAlgorithm" in code and _lib.LIB_PATH in code
    return code


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("single", [False, True])
def test_integration_stub_matches_oracle(metric, single):
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md#stub", "exec"), ns)
    ns["_lib"].kb2_last_error.restype = __import__("ctypes").c_char_p
    rng = np.random.default_rng(5)
    source = rng.standard_normal((700, 48)).astype(np.float32)
    target = None if single else rng.standard_normal((900, 48)).astype(np.float32)
    algo = ns["B200"](n_candidates=10, metric=metric)
    algo.fit(source, target)
    dist, ind = algo.kneighbors(10)
    tgt64 = (source if single else target).astype(np.float64)
    want_d, want_i = O.knn_brute(source.astype(np.float64), tgt64, 10, metric, exclude_self=single)
    O.assert_neighbors_match(dist.cpu().numpy(), ind.cpu().numpy(), want_d, want_i, 1e-5, 5e-6,
                             what=f"stub/{metric}")
