"""not gpu: the N>1 path (shard -> per-shard top-k -> all-gather -> merge) with world_size=2
over gloo on CPU.  The search and merge stages are injected (oracle arithmetic / numpy), so
what is tested is the host logic the NCCL run shares: shard bounds, global ids, the
self-exclusion offset, the packed gather layout and the merge order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kiez_b200.distributed import (RowShardComm, shard_bounds, sharded_knn_both, sharded_knn_both_rows,
                                   sharded_topk, upload_sharded)
from oracle import kiez_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _numpy_merge(g_dist, g_ind, nparts, part_stride, k, nq):
    """Reference semantics of kb2_topk_rows(nparts>1): row r = concat over parts of
    dist[p*part_stride + r*k .. +k), sorted by (value, position)."""
    gd, gi = g_dist.numpy(), g_ind.numpy()
    rows_d = np.stack([np.concatenate([gd[p * part_stride + r * k: p * part_stride + (r + 1) * k]
                                       for p in range(nparts)]) for r in range(nq)])
    rows_i = np.stack([np.concatenate([gi[p * part_stride + r * k: p * part_stride + (r + 1) * k]
                                       for p in range(nparts)]) for r in range(nq)])
    order = np.argsort(rows_d, axis=1, kind="stable")[:, :k]
    return (torch.from_numpy(np.take_along_axis(rows_d, order, 1)),
            torch.from_numpy(np.take_along_axis(rows_i, order, 1)))


def _worker(rank, world, port, q, y, k, exclude_self, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def local_search(lo, hi):
            if hi <= lo:
                return (torch.full((q.shape[0], k), float("inf"), dtype=torch.float64),
                        torch.full((q.shape[0], k), -1, dtype=torch.int64))
            kk = min(k, hi - lo)
            d = O._pairwise(q, y[lo:hi], "euclidean")
            if exclude_self:           # global column id == row id
                rows = np.arange(q.shape[0])
                inside = (rows >= lo) & (rows < hi)
                d[rows[inside], rows[inside] - lo] = np.inf
            order = np.argsort(d, axis=1, kind="stable")[:, :kk]
            dd = np.full((q.shape[0], k), np.inf)
            ii = np.full((q.shape[0], k), -1, dtype=np.int64)
            dd[:, :kk] = np.take_along_axis(d, order, 1)
            ii[:, :kk] = order + lo
            ii[np.isinf(dd)] = -1
            return torch.from_numpy(dd), torch.from_numpy(ii)

        d, i = sharded_topk(local_search, _numpy_merge, y.shape[0], k)
        out[rank] = (d.numpy(), i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("nq", "ny", "k", "exclude_self"),
                         [(37, 101, 5, False), (64, 64, 7, True), (5, 3, 2, False)])
def test_sharded_topk_world2_gloo(nq, ny, k, exclude_self):
    rng = np.random.default_rng(nq + ny)
    y = rng.standard_normal((ny, 8))
    q = y.copy() if exclude_self else rng.standard_normal((nq, 8))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, y, k, exclude_self, out), nprocs=world, join=True)
    want_d, want_i = O.knn_brute(q, y, k, exclude_self=exclude_self)
    for rank in range(world):                      # replicated result on every rank
        d, i = out[rank]
        O.assert_neighbors_match(d, i, want_d, want_i, rtol=1e-12, atol=1e-12, what=f"rank{rank}")


def test_shard_bounds_cover():
    assert [shard_bounds(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(2, 4, 3) == (2, 2)


def _upload_worker(rank, world, port, data, as_tensor, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank is given the same host matrix but may only read its own slice
        per = -(-data.shape[0] // world)
        mine = data.copy()
        mine[: rank * per] = np.nan
        mine[(rank + 1) * per:] = np.nan
        full = upload_sharded(torch.from_numpy(mine) if as_tensor else mine, "cpu")
        out[rank] = full.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("n", "dtype", "as_tensor"), [
    (101, np.float32, False), (64, np.float64, False), (7, np.float32, True), (33, np.float16, False)])
def test_upload_sharded_world2_gloo(n, dtype, as_tensor):
    """Each rank uploads ceil(n / world) rows, one all-gather completes the matrix (row counts
    that do not divide by the world size; fp64 kept, other dtypes -> fp32)."""
    data = np.random.default_rng(n).standard_normal((n, 5)).astype(dtype)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_upload_worker, args=(world, _free_port(), data, as_tensor, out), nprocs=world,
             join=True)
    want = data.astype(np.float64 if dtype == np.float64 else np.float32)
    for rank in range(world):
        assert out[rank].dtype == want.dtype and out[rank].shape == want.shape
        np.testing.assert_array_equal(out[rank], want)


class _Rows:
    """Stand-in for PreparedRows on CPU: a row range of a host matrix with its global base."""

    def __init__(self, x, base=0):
        self.x, self.base, self.n = x, base, x.shape[0]

    def rows(self, lo, hi):
        return _Rows(self.x[lo:hi], self.base + lo)


class _OracleAlgo:
    """What sharded_knn_both needs from B200: `device` and `search_both` (oracle arithmetic,
    ids global through the bases, like the CUDA path)."""

    device = torch.device("cpu")

    def search_both(self, rows, cols, k_rows, k_cols, exclude_self_rows=False):
        d_cols = O._pairwise(rows.x, cols.x, "euclidean")
        d = d_cols.copy()
        if exclude_self_rows:      # row side only (kiez's forward pass): global column id == row id
            r = np.arange(rows.n)
            c = r + rows.base - cols.base
            ok = (c >= 0) & (c < cols.n)
            d[r[ok], c[ok]] = np.inf
        def top(mat, k, base):
            kk = min(k, mat.shape[1])
            order = np.argsort(mat, axis=1, kind="stable")[:, :kk]
            dd = np.full((mat.shape[0], k), np.inf)
            ii = np.full((mat.shape[0], k), -1, dtype=np.int64)
            dd[:, :kk] = np.take_along_axis(mat, order, 1)
            ii[:, :kk] = order + base
            ii[np.isinf(dd)] = -1
            return torch.from_numpy(dd), torch.from_numpy(ii)
        return top(d, k_rows, cols.base), top(d_cols.T.copy(), k_cols, rows.base)


def _both_worker(rank, world, port, x, y, k, single, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fwd, rev = sharded_knn_both(_OracleAlgo(), _Rows(x), _Rows(y), k, k, single,
                                    merge=_numpy_merge)
        out[rank] = tuple(t.numpy() for t in (*fwd, *rev))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("nx", "ny", "k", "single"), [(40, 31, 4, False), (33, 33, 5, True),
                                                        (9, 3, 2, False)])
def test_sharded_knn_both_world2_gloo(nx, ny, k, single):
    """Dual-direction pass over column shards: row-wise lists merged across ranks, column-wise
    results complete per shard and only all-gathered (uneven shards are padded and trimmed)."""
    rng = np.random.default_rng(nx * ny)
    x = rng.standard_normal((nx, 6))
    y = x.copy() if single else rng.standard_normal((ny, 6))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_both_worker, args=(world, _free_port(), x, y, k, single, out), nprocs=world, join=True)
    k_fwd = min(k, ny - (1 if single else 0))
    want_fd, want_fi = O.knn_brute(x, y, k_fwd, exclude_self=single)
    want_rd, want_ri = O.knn_brute(y, x, min(k, nx))
    for rank in range(world):
        fd, fi, rd, ri = out[rank]
        O.assert_neighbors_match(fd[:, :k_fwd], fi[:, :k_fwd], want_fd, want_fi, 1e-12, 1e-12,
                                 what=f"fwd rank{rank}")
        O.assert_neighbors_match(rd[:, :min(k, nx)], ri[:, :min(k, nx)], want_rd, want_ri, 1e-12,
                                 1e-12, what=f"rev rank{rank}")


EMPTY = np.int64(0xFF800000FFFFFFFF - (1 << 64))


def _pack(key, row):
    """select.cuh pack_entry on the host: order-preserving fp32 key bits << 32 | row."""
    u = np.asarray(key, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = np.where((u >> np.uint64(31)) == 1, u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
    return ((u << np.uint64(32)) | np.asarray(row, dtype=np.uint64)).view(np.int64)


def _numpy_kth(gathered, nparts, m, width, kth, tau):
    """kb2_kth_key on CPU tensors: tau = min(tau, kth smallest of the gathered keys)."""
    g = gathered.numpy().transpose(1, 0, 2).reshape(m, nparts * width)
    tau.copy_(torch.minimum(tau, torch.from_numpy(np.sort(g, axis=1)[:, kth - 1].copy())))


class _RowShardOracleAlgo:
    """`B200.search_both(comm=...)` restated with numpy on the host: the same protocol -- sample
    thresholds agreed with `kth_over_ranks`, emits below the threshold, heads to the column
    owners, merge, exact finish -- over float32 keys, so that the gloo test drives every
    collective of `RowShardComm` exactly as the CUDA path does."""

    device = torch.device("cpu")

    def search_both(self, rows, cols, k_rows, k_cols, exclude_self_rows=False, comm=None):
        cap = max(k_rows, k_cols) + 2
        d2 = O._pairwise(rows.x, cols.x, "sqeuclidean")                 # [rows][cols]
        key = d2.astype(np.float32)
        # row side: complete per rank
        d = np.sqrt(d2)
        if exclude_self_rows:
            r = np.arange(rows.n)
            c = r + rows.base - cols.base
            ok = (c >= 0) & (c < cols.n)
            d[r[ok], c[ok]] = np.inf
        order = np.argsort(d, axis=1, kind="stable")[:, :k_rows]
        fwd = (torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(order + cols.base))
        # thresholds: the sample comes from ALL rows; every rank searches it for its column shard
        full = comm.rows_full.x
        sample = full[:: max(1, full.shape[0] // max(cap, full.shape[0] // 3))]
        c0s, c1s, _per = comm.column_shard(cols.n)
        s_key = O._pairwise(cols.x[c0s:c1s], sample, "sqeuclidean").astype(np.float32)
        tau_loc = np.sort(s_key, axis=1)[:, min(cap, s_key.shape[1]) - 1] if c1s > c0s else np.zeros(0, np.float32)
        tau = comm.gather_columns(torch.from_numpy(np.ascontiguousarray(tau_loc, dtype=np.float32)), cols.n)
        # two row segments: emits strictly below the threshold, the best cap of them are the
        # column's head; between the segments the ranks agree on the cap-th best key so far
        bound = tau.numpy().copy()
        kept = [np.zeros(0, dtype=np.int64) for _ in range(cols.n)]
        half = rows.n // 2
        for seg_no, (lo, hi) in enumerate(((0, half), (half, rows.n))):
            keys = np.full((cols.n, cap), np.inf, dtype=np.float32)
            for c in range(cols.n):
                hit = lo + np.flatnonzero(key[lo:hi, c] < bound[c])
                cand = np.concatenate([kept[c], hit])
                kept[c] = cand[np.argsort(key[cand, c], kind="stable")][:cap]
                if len(cand) >= cap:
                    bound[c] = key[kept[c][-1], c]
                keys[c, : len(kept[c])] = key[kept[c], c]
            if seg_no == 0:
                t = comm.kth_over_ranks(torch.from_numpy(keys), cap, tau=torch.from_numpy(bound.copy()))
                bound = t.numpy().copy()
        heads = np.full((cols.n, cap), EMPTY, dtype=np.int64)
        for c in range(cols.n):
            heads[c, : len(kept[c])] = _pack(key[kept[c], c], kept[c] + rows.base)
        tau = comm.min_(torch.from_numpy(bound))
        recv, c0, c1 = comm.columns_to_owners(torch.from_numpy(heads), int(EMPTY))
        merged = recv.numpy().transpose(1, 0, 2)[: c1 - c0].reshape(c1 - c0, -1)
        merged = np.sort(merged.view(np.uint64), axis=1)[:, :cap]       # packed order = key order
        ids = (merged & np.uint64(0xFFFFFFFF)).astype(np.int64)
        ids[ids == 0xFFFFFFFF] = -1
        # exact finish of the owner's columns against ALL rows
        rd = np.full((c1 - c0, k_cols), np.inf)
        ri = np.full((c1 - c0, k_cols), -1, dtype=np.int64)
        for j in range(c1 - c0):
            cand = ids[j][ids[j] >= 0]
            dist_j = np.sqrt(((full[cand] - cols.x[c0 + j]) ** 2).sum(axis=1))
            o = np.lexsort((cand, dist_j))[:k_cols]
            rd[j, : len(o)], ri[j, : len(o)] = dist_j[o], cand[o]
        return fwd, (torch.from_numpy(rd), torch.from_numpy(ri))


def _rows_worker(rank, world, port, x, y, k, single, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = _Rows(x)
        comm = RowShardComm(rows, kth=_numpy_kth)
        fwd, rev = sharded_knn_both_rows(_RowShardOracleAlgo(), rows, _Rows(y), k, k, single,
                                         comm=comm)
        out[rank] = tuple(t.numpy() for t in (*fwd, *rev))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("nx", "ny", "k", "single", "world"), [
    (40, 31, 4, False, 2), (33, 33, 5, True, 2), (50, 37, 3, False, 3), (64, 10, 4, False, 2)])
def test_sharded_knn_both_rows_gloo(nx, ny, k, single, world):
    """Row-sharded dual-direction pass: thresholds agreed over the ranks (all-gather + kth),
    column heads sent to the column owners (all-to-all, ragged column shards padded), both
    results replicated with one all-gather each."""
    rng = np.random.default_rng(nx * ny + world)
    x = rng.standard_normal((nx, 6)).astype(np.float32).astype(np.float64)
    y = x.copy() if single else rng.standard_normal((ny, 6)).astype(np.float32).astype(np.float64)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rows_worker, args=(world, _free_port(), x, y, k, single, out), nprocs=world, join=True)
    want_fd, want_fi = O.knn_brute(x, y, k, exclude_self=single)
    want_rd, want_ri = O.knn_brute(y, x, k)
    for rank in range(world):
        fd, fi, rd, ri = out[rank]
        assert fd.shape == (nx, k) and rd.shape == (y.shape[0], k)
        O.assert_neighbors_match(fd, fi, want_fd, want_fi, 1e-9, 1e-9, what=f"rows fwd rank{rank}")
        O.assert_neighbors_match(rd, ri, want_rd, want_ri, 1e-9, 1e-9, what=f"rows rev rank{rank}")


def _dsl_worker(rank, world, port, source, target, fwd_i, rev_i, k, squared, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kiez_b200.distributed import all_gather_vector, dsl_transform_sharded

        m, n = target.shape[0], source.shape[0]
        # _fit sharded over the target rows: per-target scalars all-gathered
        t0, t1 = shard_bounds(m, world, rank)
        cent = source[rev_i[t0:t1]].mean(axis=1)
        d2c = all_gather_vector(torch.from_numpy(((target[t0:t1] - cent) ** 2).sum(axis=1)), m).numpy()

        def raw_fn(lo, hi):
            q, ind = source[lo:hi], fwd_i[lo:hi]
            sq = ((q[:, None, :] - target[ind]) ** 2).sum(axis=2)
            qc = ((q - target[ind].mean(axis=1)) ** 2).sum(axis=1)
            raw = sq - qc[:, None] - d2c[ind]
            gmin = torch.tensor([raw.min() if raw.size else np.inf], dtype=torch.float64)
            return raw, ind, gmin

        def finish_fn(raw, ind, gmin):
            v = raw + (-float(gmin) if float(gmin) < 0 else 0.0)
            v = v if squared else np.sqrt(v)
            d, i = O.sort_topk(v, ind, k)
            return torch.from_numpy(d), torch.from_numpy(i)

        d, i = dsl_transform_sharded(n, raw_fn, finish_fn)
        out[rank] = (d.numpy(), i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("n", "m", "world", "squared"), [(41, 30, 2, False), (7, 9, 3, True)])
def test_dsl_global_minimum_all_reduce_gloo(n, m, world, squared):
    """DisSimLocal with sharded rows: the shift is the minimum of the WHOLE (n, c) matrix
    (dis_sim.py:171-173), so the ranks all-reduce(MIN) their local minima between the two stages;
    the result must equal the oracle's single-process transform on every rank."""
    rng = np.random.default_rng(n * m)
    source, target = rng.standard_normal((n, 5)), rng.standard_normal((m, 5))
    c, k = 4, 3
    _fd, fwd_i = O.knn_brute(source, target, c)
    _rd, rev_i = O.knn_brute(target, source, c)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dsl_worker, args=(world, _free_port(), source, target, fwd_i, rev_i, k, squared, out),
             nprocs=world, join=True)
    want = O.dsl_transform(fwd_i, source, target, O.dsl_fit(rev_i, source, target)[1], squared=squared)
    want_d, want_i = O.sort_topk(want, fwd_i, k)
    local_mins = []
    for rank in range(world):
        lo, hi = shard_bounds(n, world, rank)
        local_mins.append(want[lo:hi].min() if hi > lo else np.inf)
        d, i = out[rank]
        O.assert_neighbors_match(d, i, want_d, want_i, 1e-9, 1e-9, what=f"dsl rank{rank}")
    assert len(set(np.round(local_mins, 9))) > 1      # the shards really disagree on their minimum
