"""-m gpu: host inputs take the overlapped upload (kiez_b200/upload.py): row chunks through a
pinned staging ring on a background thread, the index build and the dual-direction pass consume
them as they arrive.  Results must equal those of the same fit from device tensors (the plain
path) and the oracle; pageable numpy, pinned numpy views and CPU tensors are covered, and the
shapes that are NOT eligible (float64, small, single source) still work."""
import numpy as np
import pytest

from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RTOL, ATOL = 1e-5, 5e-6


def _data(n, m, d, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, d)).astype(np.float32),
            rng.standard_normal((m, d)).astype(np.float32))


def _fit_predict(source, target, hubness, fused, c=10, k=5, sample_chunk=None):
    from kiez_b200 import B200, Kiez

    algo = B200(n_candidates=c, fused=fused)
    algo.FUSED_SEGMENT_MIN_ROWS = 4096
    if sample_chunk:
        algo.SAMPLE_CHUNK_ROWS = sample_chunk    # threshold search in chunks of the arriving target
    inst = Kiez(n_candidates=c, algorithm=algo, hubness=hubness)
    inst.fit(source, target)
    return inst, inst.kneighbors(k)


@pytest.mark.parametrize("kind", ["pageable", "pinned", "cpu_tensor"])
@pytest.mark.parametrize("fused", [True, False])
def test_overlapped_upload_matches_device_inputs(kind, fused):
    n, m, d = 70000, 66000, 128                      # 36 MB / 34 MB: two chunks each
    source, target = _data(n, m, d, 5)
    if kind == "pinned":
        src = torch.from_numpy(source).pin_memory().numpy()
        tgt = torch.from_numpy(target).pin_memory().numpy()
    elif kind == "cpu_tensor":
        src, tgt = torch.from_numpy(source), torch.from_numpy(target)
    else:
        src, tgt = source, target
    inst, (dist, ind) = _fit_predict(src, tgt, "CSLS", fused,
                                     sample_chunk=16384 if kind == "pageable" else None)
    algo = inst.algorithm
    assert algo._uploader is not None and len(algo._uploader.jobs) == 3     # target, sample, source
    assert algo._prepared[id(src)]._pending is None                         # everything consumed
    if kind != "cpu_tensor":
        assert isinstance(dist, np.ndarray) and ind.dtype == np.int64
    ref_inst, (ref_d, ref_i) = _fit_predict(torch.from_numpy(source).cuda(),
                                            torch.from_numpy(target).cuda(), "CSLS", fused)
    assert ref_inst.algorithm._uploader is None
    # the centring vector differs by sampling only for device input of the same size: identical here
    np.testing.assert_array_equal(np.asarray(torch.as_tensor(ind).cpu()), ref_i.cpu().numpy())
    np.testing.assert_allclose(np.asarray(torch.as_tensor(dist).cpu()), ref_d.cpu().numpy(),
                               rtol=1e-12, atol=0)
    # and the oracle on a row sample (forward kNN, before the rescale)
    rows = np.random.default_rng(1).choice(n, 200, replace=False)
    fd, fi = algo.kneighbors(k=10)
    want_d, want_i = O.knn_sklearn(source[rows].astype(np.float64), target.astype(np.float64), 10)
    O.assert_neighbors_match(fd[rows].cpu().numpy(), fi[rows].cpu().numpy(), want_d, want_i,
                             RTOL, ATOL, what=f"upload {kind} fused={fused}")


def test_overlapped_upload_other_entry_points():
    from kiez_b200 import B200, Kiez

    n, m, d = 40000, 30000, 96
    source, target = _data(n, m, d, 6)
    s64, t64 = source.astype(np.float64), target.astype(np.float64)
    rows = np.random.default_rng(2).choice(n, 150, replace=False)
    # no hubness reduction: only the target is fitted and uploaded
    inst = Kiez(n_candidates=10, algorithm=B200(n_candidates=10), hubness=None)
    inst.fit(source, target)
    assert len(inst.algorithm._uploader.jobs) == 1
    dist, ind = inst.kneighbors(10)
    want_d, want_i = O.knn_sklearn(s64[rows], t64, 10)
    O.assert_neighbors_match(dist[rows], ind[rows], want_d, want_i, RTOL, ATOL, what="no hubness")
    # DisSimLocal reads the raw rows of both matrices after the searches
    inst = Kiez(n_candidates=10, algorithm=B200(n_candidates=10), hubness="DisSimLocal")
    inst.fit(source, target)
    assert len(inst.algorithm._uploader.jobs) == 3
    dist, ind = inst.kneighbors(5)
    want_d, want_i = O.kiez_kneighbors(s64, t64, hubness="dsl", n_candidates=10, k=5,
                                       knn=O.knn_sklearn)
    O.assert_neighbors_match(dist, ind, want_d, want_i, RTOL, ATOL, what="dsl")
    # single source: one matrix, one job; float64 and small inputs are not eligible (plain path)
    inst = Kiez(n_candidates=10, algorithm=B200(n_candidates=10), hubness="CSLS")
    inst.fit(source)
    assert len(inst.algorithm._uploader.jobs) == 1
    dist, ind = inst.kneighbors(5)
    want_fd, want_fi = O.knn_sklearn(s64[rows], s64, 11)
    got_fd, got_fi = inst.algorithm.kneighbors(k=10)
    O.assert_neighbors_match(got_fd[rows].cpu().numpy(), got_fi[rows].cpu().numpy(), want_fd[:, 1:],
                             want_fi[:, 1:], RTOL, ATOL, what="single source")
    inst.fit(source[:3000])
    assert inst.algorithm._uploader is None
    inst.fit(s64[:3000], t64[:2500])
    assert inst.algorithm._uploader is None


def test_overlapped_upload_can_be_disabled(monkeypatch):
    monkeypatch.setenv("KB2_ASYNC_UPLOAD", "0")
    source, target = _data(40000, 30000, 96, 7)
    inst, (dist, ind) = _fit_predict(source, target, "CSLS", True)
    assert inst.algorithm._uploader is None and isinstance(dist, np.ndarray)
