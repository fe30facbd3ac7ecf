"""-m gpu (needs >= 2 GPUs, skipped otherwise): the NCCL paths end to end -- one-direction passes
with the index rows sharded over 2 ranks (per-shard search, all-gather, GPU merge kernel), the
dual-direction pass with the source rows sharded (thresholds agreed between row segments,
column heads sent to the column owners) or, for small problems, the target columns sharded --
must reproduce the oracle on every rank."""
import os
import socket

import numpy as np
import pytest

from oracle import kiez_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, source, target, hub, kw, c, k, out, fused=False):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from kiez_b200 import B200, Kiez

        inst = Kiez(n_candidates=c,
                    algorithm=B200(n_candidates=c, distributed=True, fused=fused), hubness=hub,
                    hubness_kwargs=dict(kw))
        inst.fit(source, target)
        d, i = inst.kneighbors(k)
        out[rank] = (np.asarray(d), np.asarray(i))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize(("hub", "kw", "label"), [
    ("CSLS", {}, "csls"), (None, {}, "no"), ("MutualProximity", {"method": "normal"}, "mp_gaussian"),
    ("DisSimLocal", {}, "dsl")])
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_two_gpu_matches_oracle(hub, kw, label, single, fused):
    import torch.multiprocessing as mp

    rng = np.random.default_rng(31)
    source = rng.standard_normal((1500, 64)).astype(np.float32)
    target = None if single else rng.standard_normal((2100, 64)).astype(np.float32)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), source, target, hub, kw, 16, 8, out, fused), nprocs=2,
             join=True)
    want_d, want_i = O.kiez_kneighbors(source.astype(np.float64),
                                       None if single else target.astype(np.float64),
                                       hubness=label, n_candidates=16, k=8)
    for rank in range(2):
        d, i = out[rank]
        O.assert_neighbors_match(d, i, want_d, want_i, 1e-5, 5e-6, what=f"{label} rank{rank}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_two_gpu_sharded_upload(fused, dtype):
    """Host inputs of >= 4096 rows: every rank uploads its slice, NCCL all-gathers the matrix
    (B200._upload_sharded); row counts that do not divide by the world size."""
    import torch.multiprocessing as mp

    rng = np.random.default_rng(32)
    source = rng.standard_normal((4099, 32)).astype(dtype)
    target = rng.standard_normal((5001, 32)).astype(dtype)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), source, target, "CSLS", {}, 10, 5, out, fused),
             nprocs=2, join=True)
    want_d, want_i = O.kiez_kneighbors(source.astype(np.float64), target.astype(np.float64),
                                       hubness="csls", n_candidates=10, k=5)
    for rank in range(2):
        d, i = out[rank]
        O.assert_neighbors_match(d, i, want_d, want_i, 1e-5, 5e-6, what=f"sharded upload rank{rank}")


def _rows_worker(rank, world, port, source, target, c, k, knobs, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from kiez_b200 import B200, Kiez

        algo = B200(n_candidates=c, distributed=True, fused=True, shard_mode=knobs.get("mode", "rows"),
                    precision=knobs.get("precision", "auto"))
        algo.FUSED_SEGMENT_MIN_ROWS = knobs.get("min_rows", 512)
        algo.FUSED_COL_CAP = knobs.get("col_cap", 512)
        inst = Kiez(n_candidates=c, algorithm=algo, hubness="CSLS")
        inst.fit(source, target)
        d, i = inst.kneighbors(k)
        fd, fi = algo.kneighbors(k=c)
        rd, ri = inst.hubness.r_dist_train_, inst.hubness.r_ind_train_
        out[rank] = tuple(np.asarray(torch.as_tensor(t).cpu()) for t in (d, i, fd, fi, rd, ri)) + (
            dict(algo.search_stats), getattr(algo, "_fused_stats", None))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize(("label", "data", "single", "knobs"), [
    ("rows", "gauss", False, {}),
    ("rows-single-source", "gauss", True, {}),
    ("rows-tf32x3", "gauss", False, {"precision": "tf32x3"}),
    ("rows-overflow", "gauss", False, {"col_cap": 24}),
    ("rows-hubby", "hubby", False, {}),
    ("cols", "gauss", False, {"mode": "cols"}),
])
def test_two_gpu_dual_direction_pass(label, data, single, knobs):
    """The row-sharded dual-direction pass with several row segments per rank (thresholds
    exchanged between them), ragged shards, column-buffer overflow (re-searched by the column
    owner), clustered data (the probe verdict is agreed across ranks), and the column-shard
    fallback: forward kNN, reverse kNN and the CSLS result equal the oracle on both ranks."""
    import torch.multiprocessing as mp

    rng = np.random.default_rng(41)
    n, m, d, c, k = 9001, 7003, 64, 10, 5

    def synth(rows):
        x = rng.standard_normal((rows, d))
        if data == "hubby":
            cent = np.random.default_rng(5).standard_normal((4, d))
            x = 0.25 * x + cent[rng.integers(0, 4, rows)]
            x /= np.linalg.norm(x, axis=1, keepdims=True)
        return x.astype(np.float32)

    source = synth(n)
    target = None if single else synth(m)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rows_worker, args=(2, _free_port(), source, target, c, k, knobs, out), nprocs=2,
             join=True)
    s64 = source.astype(np.float64)
    t64 = s64 if single else target.astype(np.float64)
    want_fd, want_fi = O.knn_brute(s64, t64, c, exclude_self=single)
    want_rd, want_ri = O.knn_brute(t64, s64, c)
    want_d, want_i = O.kiez_kneighbors(s64, None if single else t64, hubness="csls",
                                       n_candidates=c, k=k)
    for rank in range(2):
        d_, i_, fd, fi, rd, ri, stats, fstats = out[rank]
        O.assert_neighbors_match(fd, fi, want_fd, want_fi, 1e-5, 5e-6, what=f"{label} fwd rank{rank}")
        O.assert_neighbors_match(rd, ri, want_rd, want_ri, 1e-5, 5e-6, what=f"{label} rev rank{rank}")
        O.assert_neighbors_match(d_, i_, want_d, want_i, 1e-5, 5e-6, what=f"{label} csls rank{rank}")
        print(label, "rank", rank, stats, fstats)
    assert out[0][6].get("screen_probe_unverified") == out[1][6].get("screen_probe_unverified")
