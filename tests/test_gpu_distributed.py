"""-m gpu (needs >= 2 GPUs, skipped otherwise): the NCCL path end to end -- index rows sharded
over 2 ranks, per-shard search, all-gather, GPU merge kernel -- must reproduce the single-GPU
result and the oracle."""
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
