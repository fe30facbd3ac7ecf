"""Multi-GPU exact kNN: one process per GPU (torch.distributed, NCCL over NVLink).

One-direction passes: the index side is sharded by rows over the ranks (forward pass:
target rows; reverse pass: source rows -- SURVEY.md section 8e); queries are replicated.
Every rank searches its shard for all queries (candidate search + exact finish,
ids made global with the shard's base), the per-shard top-k lists are exchanged
with ONE all-gather, and a GPU merge kernel (kb2_topk_rows, nparts = world size)
keeps the k best per query.  The result is replicated on every rank, so the
rescaling that follows needs no further communication.

Dual-direction pass (kiez's reverse + forward kNN from one contraction, the default at
BASELINE.json's metric config): the SOURCE rows are sharded (`sharded_knn_both_rows` +
`RowShardComm`); every rank contracts its rows with all targets, the ranks agree on the
per-target thresholds between row segments and each rank finishes a shard of the targets.
`sharded_knn_both` (target shards, per-rank thresholds) remains for problems too small to
give every rank a few row tiles.

`shard_bounds`, the gather layout and the merge are backend-agnostic
(`sharded_topk`), which is how the world_size-2 gloo tests exercise them on CPU.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first n_rows % world ranks get one more)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_slice(n_rows: int, group=None) -> Tuple[int, int, int]:
    """(lo, hi, per): the rows this rank uploads of a host matrix every rank holds."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_rows // world)
    return min(n_rows, rank * per), min(n_rows, (rank + 1) * per), per


def upload_sharded(data, device, group=None, mine=None) -> torch.Tensor:
    """A host matrix every rank holds (numpy or CPU tensor, fp32 / fp64 kept, anything else ->
    fp32) -> the full matrix on `device`, moving only 1/world of it over this rank's
    host-to-device link: rank r uploads rows [r * per, (r + 1) * per), per = ceil(n / world), and
    ONE all-gather (NVLink under NCCL) completes the matrix.  Backend-agnostic: the gloo tests
    run it with device = "cpu"."""
    import numpy as np

    world = dist.get_world_size(group)
    n, d = data.shape
    lo, hi, per = shard_slice(n, group)
    if isinstance(data, np.ndarray):
        dtype = torch.float64 if data.dtype == np.float64 else torch.float32
    else:
        dtype = torch.float64 if data.dtype == torch.float64 else torch.float32
    full = torch.empty((world * per, d), dtype=dtype, device=device)
    # the last rank's slice may be short: the padding rows are never read (full[:n])
    padded = torch.zeros((per, d), dtype=dtype, device=device)
    if mine is not None:         # the slice is on the device already (overlapped upload, fp32)
        padded[: hi - lo].copy_(mine)
    elif isinstance(data, np.ndarray):
        padded[: hi - lo].copy_(torch.from_numpy(np.ascontiguousarray(
            data[lo:hi], dtype=np.float64 if dtype == torch.float64 else np.float32)),
            non_blocking=True)
    else:
        padded[: hi - lo].copy_(data[lo:hi].to(dtype).contiguous(), non_blocking=True)
    dist.all_gather_into_tensor(full, padded, group=group)
    return full[:n]


def sharded_topk(local_search: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]],
                 merge: Callable[[torch.Tensor, torch.Tensor, int, int, int],
                                 Tuple[torch.Tensor, torch.Tensor]],
                 n_index: int, k: int, group=None):
    """Generic shard -> all-gather -> merge.

    local_search(lo, hi) -> (dist (nq,k) float64, ind (nq,k) int64 with GLOBAL ids; slots a
    small shard cannot fill hold +inf / -1).  merge(gathered_dist, gathered_ind, nparts,
    part_stride, k) -> (dist, ind), where gathered_* are views of ONE packed buffer laid out
    [rank][dist bits | ind][nq][k].
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_index, world, rank)
    d, i = local_search(lo, hi)
    nq = d.shape[0]
    packed = torch.empty(2 * nq * k, dtype=torch.int64, device=d.device)
    packed[: nq * k] = d.contiguous().view(torch.int64).reshape(-1)
    packed[nq * k:] = i.reshape(-1)
    gathered = torch.empty(world * 2 * nq * k, dtype=torch.int64, device=d.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    g_dist = gathered.view(torch.float64)
    g_ind = gathered[nq * k:]
    return merge(g_dist, g_ind, world, 2 * nq * k, k, nq)


def device_merge(g_dist, g_ind, nparts, part_stride, k, nq):
    """kb2_topk_rows over the gathered per-shard lists (the multi-GPU merge kernel)."""
    from . import _lib as lib

    dev = g_dist.device
    with torch.cuda.device(dev):
        od = torch.empty((nq, k), dtype=torch.float64, device=dev)
        oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq:
            lib.call("kb2_topk_rows", lib.ptr(g_dist), lib.ptr(g_ind), nq, k, nparts, part_stride,
                     k, lib.ptr(od), lib.ptr(oi), lib.stream_ptr())
    return od, oi


def sharded_knn(algo, q, index, k: int, exclude_self: bool, group=None):
    """`B200._kneighbors` in distributed mode: (dist, ind) replicated on every rank."""

    def local_search(lo, hi):
        if hi <= lo:   # more ranks than index rows
            d = torch.full((q.n, k), float("inf"), dtype=torch.float64, device=algo.device)
            i = torch.full((q.n, k), -1, dtype=torch.int64, device=algo.device)
            return d, i
        return algo.search(q, index.rows(lo, hi), k, exclude_self=exclude_self)

    return sharded_topk(local_search, device_merge, index.n, k, group=group)


def sharded_knn_both(algo, rows, cols, k_fwd: int, k_rev: int, exclude_self_rows: bool, group=None,
                      merge=None):
    """Distributed dual-direction pass.  Each rank owns a contiguous shard of the COLUMNS
    (targets) and sees all rows (sources): its pass yields the row-wise lists over its shard
    (merged across ranks like `sharded_knn`) and the COMPLETE column-wise result for its own
    columns (every row was visited locally), which only needs an all-gather to be replicated.
    Returns ((fwd_dist, fwd_ind), (rev_dist, rev_ind)), identical on every rank.
    `merge` defaults to the GPU merge kernel (`device_merge`); the gloo tests inject numpy."""
    merge = device_merge if merge is None else merge
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(cols.n, world, rank)
    dev = algo.device
    if hi > lo:
        (fd, fi), (rd, ri) = algo.search_both(rows, cols.rows(lo, hi), k_fwd, k_rev,
                                              exclude_self_rows=exclude_self_rows)
    else:
        fd = torch.full((rows.n, k_fwd), float("inf"), dtype=torch.float64, device=dev)
        fi = torch.full((rows.n, k_fwd), -1, dtype=torch.int64, device=dev)
        rd = torch.empty((0, k_rev), dtype=torch.float64, device=dev)
        ri = torch.empty((0, k_rev), dtype=torch.int64, device=dev)
    fwd = sharded_topk(lambda _lo, _hi: (fd, fi), merge, cols.n, k_fwd, group=group)
    # reverse: pad every shard to the largest shard, one packed all-gather, trim
    per = -(-cols.n // world)
    packed = torch.zeros(2 * per * k_rev, dtype=torch.int64, device=dev)
    packed[: (hi - lo) * k_rev] = rd.contiguous().view(torch.int64).reshape(-1)
    packed[per * k_rev: per * k_rev + (hi - lo) * k_rev] = ri.reshape(-1)
    gathered = torch.empty(world * 2 * per * k_rev, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    g = gathered.view(world, 2, per, k_rev)
    rev_d = torch.empty((cols.n, k_rev), dtype=torch.float64, device=dev)
    rev_i = torch.empty((cols.n, k_rev), dtype=torch.int64, device=dev)
    for r in range(world):
        rlo, rhi = shard_bounds(cols.n, world, r)
        rev_d[rlo:rhi] = g[r, 0, : rhi - rlo].view(torch.float64)
        rev_i[rlo:rhi] = g[r, 1, : rhi - rlo]
    return fwd, (rev_d, rev_i)


def all_gather_blocks(d, i, total: int, per: int, bounds, group=None):
    """Per-rank result blocks (rows bounds(r) of the full result, at most `per` each; d float64,
    i int64, same shape) -> the full (total, k) result on every rank: one packed all-gather
    (dist bits | ind)."""
    world = dist.get_world_size(group)
    k = d.shape[1]
    n_loc = d.shape[0]
    packed = torch.zeros(2 * per * k, dtype=torch.int64, device=d.device)
    packed[: n_loc * k] = d.contiguous().view(torch.int64).reshape(-1)
    packed[per * k: per * k + n_loc * k] = i.reshape(-1)
    gathered = torch.empty(world * 2 * per * k, dtype=torch.int64, device=d.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    g = gathered.view(world, 2, per, k)
    out_d = torch.empty((total, k), dtype=torch.float64, device=d.device)
    out_i = torch.empty((total, k), dtype=torch.int64, device=d.device)
    for r in range(world):
        lo, hi = bounds(r)
        out_d[lo:hi] = g[r, 0, : hi - lo].view(torch.float64)
        out_i[lo:hi] = g[r, 1, : hi - lo]
    return out_d, out_i


def dsl_transform_sharded(n: int, raw_fn, finish_fn, group=None):
    """DisSimLocal.transform with the QUERY rows sharded (dis_sim.py:139-181): rank r computes the
    raw values of its rows -- raw_fn(lo, hi) -> (raw, ind, local_min) --, the ranks agree on the
    GLOBAL minimum of the whole (n, c) matrix with one all-reduce(MIN) (dis_sim.py:171-173 shifts
    by it), finish_fn(raw, ind, global_min) -> (dist, ind) of the own rows, and one all-gather
    replicates the result.  Backend-agnostic (the gloo test injects numpy)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    raw, ind, gmin = raw_fn(lo, hi)
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN, group=group)
    d, i = finish_fn(raw, ind, gmin)
    return all_gather_blocks(d, i, n, -(-n // world), lambda r: shard_bounds(n, world, r), group)


def all_gather_vector(x_local, total: int, group=None):
    """Concatenate per-rank 1-d float64 shards (rank r holds shard_bounds(total, world, r))."""
    world = dist.get_world_size(group)
    per = -(-total // world)
    mine = torch.zeros(per, dtype=x_local.dtype, device=x_local.device)
    mine[: x_local.numel()] = x_local
    gathered = torch.empty(world * per, dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    out = torch.empty(total, dtype=x_local.dtype, device=x_local.device)
    for r in range(world):
        lo, hi = shard_bounds(total, world, r)
        out[lo:hi] = gathered[r * per: r * per + hi - lo]
    return out


class RowShardComm:
    """The collectives of the row-sharded dual-direction pass (`B200.search_both(comm=...)`).

    Every rank runs the pass over ITS rows against all columns.  Row-wise results are then
    complete per rank; the per-column state has to agree across ranks:
      * thresholds: after the sample search and between row segments every rank contributes its
        best keys per column, `kth_over_ranks` all-gathers them and keeps the kth best of the
        union -- the ranks continue with ONE threshold per column, so a column receives as
        many emits in total as in a single-GPU run (with per-rank thresholds it would receive
        that many PER RANK, which is what capped the column-shard scheme at 0.66 efficiency
        on 8 GPUs);
      * result: after the last segment `columns_to_owners` sends, with one all-to-all, the head
        of every column (each rank's best `cap` rows) to the rank that owns the column; the
        owner merges the `world` heads and runs the exact finish for its columns only.
    `kth` is injectable so that the world_size-2 gloo tests can run the layer on CPU tensors."""

    def __init__(self, rows_full, group=None, kth=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows_full = rows_full          # every rank holds all rows (needed by the owners' finish)
        self.n_rows = rows_full.n
        self._kth = kth if kth is not None else self._device_kth

    @staticmethod
    def _device_kth(gathered, nparts, m, width, kth, tau):
        from . import _lib as lib

        with torch.cuda.device(tau.device):
            lib.call("kb2_kth_key", lib.ptr(gathered), nparts, m * width, m, width, kth, lib.ptr(tau),
                     lib.stream_ptr())

    def kth_over_ranks(self, keys, kth, tau=None):
        """keys [m][width] fp32 (this rank's best keys per column, +inf padded) -> tau [m] =
        min(tau, kth smallest over the ranks' keys).  One all-gather of m * width * 4 bytes."""
        m, width = keys.shape
        gathered = torch.empty((self.world, m, width), dtype=torch.float32, device=keys.device)
        dist.all_gather_into_tensor(gathered.view(-1), keys.contiguous().view(-1), group=self.group)
        if tau is None:
            tau = torch.full((m,), float("inf"), dtype=torch.float32, device=keys.device)
        self._kth(gathered, self.world, m, width, kth, tau)
        return tau

    def max_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def min_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t

    def max_scalar(self, x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=self.rows_full.raw.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def column_shard(self, m: int) -> Tuple[int, int, int]:
        """(c0, c1, per): rank r owns the columns [r * per, min(m, (r + 1) * per))."""
        per = -(-m // self.world)
        return min(m, self.rank * per), min(m, (self.rank + 1) * per), per

    def gather_columns(self, x_local, m: int):
        """Per-column values of this rank's column shard (`column_shard(m)`) -> all m on every rank."""
        _c0, _c1, per = self.column_shard(m)
        mine = torch.zeros(per, dtype=x_local.dtype, device=x_local.device)
        mine[: x_local.numel()] = x_local
        gathered = torch.empty(self.world * per, dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        return gathered[:m].contiguous()

    def columns_to_owners(self, heads, pad_value):
        """heads [m][cap] int64 (this rank's best rows per column) -> (recv [world][per][cap]:
        the heads of this rank's column shard from every rank, c0, c1).  One all-to-all."""
        m, cap = heads.shape
        c0, c1, per = self.column_shard(m)
        send = heads
        if self.world * per != m:
            send = torch.full((self.world * per, cap), pad_value, dtype=heads.dtype, device=heads.device)
            send[:m] = heads
        recv = torch.empty((self.world, per, cap), dtype=heads.dtype, device=heads.device)
        dist.all_to_all_single(recv.view(-1), send.contiguous().view(-1), group=self.group)
        return recv, c0, c1

    def gather_blocks(self, d, i, total: int, per: int, bounds):
        return all_gather_blocks(d, i, total, per, bounds, self.group)


def sharded_knn_both_rows(algo, rows, cols, k_fwd: int, k_rev: int, exclude_self_rows: bool,
                          group=None, comm=None):
    """Distributed dual-direction pass with the ROWS (sources) sharded: rank r runs
    `search_both` over its contiguous row shard against all columns with `RowShardComm`
    agreeing the column state across ranks; the forward result of its rows and the reverse result
    of its column shard are then replicated with one packed all-gather each.
    Returns ((fwd_dist, fwd_ind), (rev_dist, rev_ind)), identical on every rank."""
    comm = RowShardComm(rows, group) if comm is None else comm
    world, rank = comm.world, comm.rank
    lo, hi = shard_bounds(rows.n, world, rank)
    (fd, fi), (rd, ri) = algo.search_both(rows.rows(lo, hi), cols, k_fwd, k_rev,
                                          exclude_self_rows=exclude_self_rows, comm=comm)
    per_rows = -(-rows.n // world)
    fwd = comm.gather_blocks(fd, fi, rows.n, per_rows, lambda r: shard_bounds(rows.n, world, r))
    per_cols = -(-cols.n // world)
    rev = comm.gather_blocks(rd, ri, cols.n, per_cols,
                             lambda r: (min(cols.n, r * per_cols), min(cols.n, (r + 1) * per_cols)))
    return fwd, rev
