"""-m gpu: from_openea(device=...) returns the same rows as the host path, resident on the
device, and feeds Kiez.fit directly."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = os.path.join(ROOT, "tests", "golden", "openea_small")


def test_from_openea_on_device_feeds_kiez():
    from kiez_b200 import Kiez, hits
    from kiez_b200.io import from_openea

    host = from_openea(os.path.join(SMALL, "emb"), os.path.join(SMALL, "kg"))
    dev = from_openea(os.path.join(SMALL, "emb"), os.path.join(SMALL, "kg"), device="cuda")
    assert dev[0].is_cuda and dev[1].is_cuda
    np.testing.assert_array_equal(dev[0].cpu().numpy(), host[0])
    np.testing.assert_array_equal(dev[1].cpu().numpy(), host[1])
    assert dev[2:] == host[2:]
    inst = Kiez(n_candidates=5, algorithm="B200", hubness="CSLS")
    inst.fit(dev[0], dev[1])
    dist, ind = inst.kneighbors(3)
    assert ind.shape == (host[0].shape[0], 3) and ind.is_cuda
    res = hits(ind, dev[4], k=[1, 3])
    assert set(res) == {1, 3}
