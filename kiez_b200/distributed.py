"""Multi-GPU exact kNN: one process per GPU (torch.distributed, NCCL over NVLink).

The index side of each pass is sharded by rows over the ranks (forward pass: target
rows; reverse pass: source rows -- SURVEY.md section 8e); queries are replicated.
Every rank searches its shard for all queries (candidate search + exact finish,
ids made global with the shard's base), the per-shard top-k lists are exchanged
with ONE all-gather, and a GPU merge kernel (kb2_topk_rows, nparts = world size)
keeps the k best per query.  The result is replicated on every rank, so the
rescaling that follows needs no further communication.

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


def upload_sharded(data, device, group=None) -> torch.Tensor:
    """A host matrix every rank holds (numpy or CPU tensor, fp32 / fp64 kept, anything else ->
    fp32) -> the full matrix on `device`, moving only 1/world of it over this rank's
    host-to-device link: rank r uploads rows [r * per, (r + 1) * per), per = ceil(n / world), and
    ONE all-gather (NVLink under NCCL) completes the matrix.  Backend-agnostic: the gloo tests
    run it with device = "cpu"."""
    import numpy as np

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, d = data.shape
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    if isinstance(data, np.ndarray):
        dtype = torch.float64 if data.dtype == np.float64 else torch.float32
        part = torch.from_numpy(np.ascontiguousarray(
            data[lo:hi], dtype=np.float64 if dtype == torch.float64 else np.float32))
    else:
        dtype = torch.float64 if data.dtype == torch.float64 else torch.float32
        part = data[lo:hi].to(dtype).contiguous()
    full = torch.empty((world * per, d), dtype=dtype, device=device)
    # the last rank's slice may be short: the padding rows are never read (full[:n])
    mine = torch.zeros((per, d), dtype=dtype, device=device)
    mine[: hi - lo].copy_(part, non_blocking=True)
    dist.all_gather_into_tensor(full, mine, group=group)
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


def sharded_knn_both_grid(algo, rows, cols, k_fwd: int, k_rev: int, exclude_self_rows: bool,
                          grid: Tuple[int, int], group=None, merge=None):
    """Dual-direction pass on an R x C grid of ranks (R * C = world size): rank (r, c) =
    divmod(rank, C) contracts row block r with column block c.  Compared with the column shards
    of `sharded_knn_both` (= grid (1, world)) every row list sees world / R times more columns
    (a shorter list fill phase per flop) and every rank finishes only n / R row lists; the price
    is a second merge: row-wise lists are merged across the C column blocks of a row block,
    column-wise lists across the R row blocks of a column block.  Two packed all-gathers, as
    before.  EXPERIMENTAL: host logic covered by the gloo tests, not yet the default (DESIGN.md
    round-2 list).  Returns ((fwd_dist, fwd_ind), (rev_dist, rev_ind)), identical on every rank."""
    merge = device_merge if merge is None else merge
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_r, n_c = grid
    if n_r * n_c != world:
        raise ValueError(f"grid {grid} does not match world size {world}")
    r, c = divmod(rank, n_c)
    per_r, per_c = -(-rows.n // n_r), -(-cols.n // n_c)
    r0, r1 = min(rows.n, r * per_r), min(rows.n, (r + 1) * per_r)
    c0, c1 = min(cols.n, c * per_c), min(cols.n, (c + 1) * per_c)
    dev = algo.device
    inf = float("inf")
    # padded per-rank results: slots a block cannot fill hold +inf / -1
    fd = torch.full((per_r, k_fwd), inf, dtype=torch.float64, device=dev)
    fi = torch.full((per_r, k_fwd), -1, dtype=torch.int64, device=dev)
    rd = torch.full((per_c, k_rev), inf, dtype=torch.float64, device=dev)
    ri = torch.full((per_c, k_rev), -1, dtype=torch.int64, device=dev)
    if r1 > r0 and c1 > c0:
        kf, kr = min(k_fwd, c1 - c0), min(k_rev, r1 - r0)
        (bfd, bfi), (brd, bri) = algo.search_both(rows.rows(r0, r1), cols.rows(c0, c1), kf, kr,
                                                  exclude_self_rows=exclude_self_rows)
        fd[: r1 - r0, :kf], fi[: r1 - r0, :kf] = bfd, bfi
        rd[: c1 - c0, :kr], ri[: c1 - c0, :kr] = brd, bri

    def gather(d, i):
        n_loc, k = d.shape
        packed = torch.empty(2 * n_loc * k, dtype=torch.int64, device=dev)
        packed[: n_loc * k] = d.contiguous().view(torch.int64).reshape(-1)
        packed[n_loc * k:] = i.reshape(-1)
        gathered = torch.empty(world * packed.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, packed, group=group)
        return gathered, packed.numel()

    # row-wise: for row block b, the parts are the C consecutive ranks b * C .. b * C + C - 1
    g, size = gather(fd, fi)
    fwd_d = torch.empty((rows.n, k_fwd), dtype=torch.float64, device=dev)
    fwd_i = torch.empty((rows.n, k_fwd), dtype=torch.int64, device=dev)
    for b in range(n_r):
        lo, hi = min(rows.n, b * per_r), min(rows.n, (b + 1) * per_r)
        if hi <= lo:
            continue
        base = b * n_c * size
        md, mi = merge(g[base:].view(torch.float64), g[base + per_r * k_fwd:], n_c, size, k_fwd, per_r)
        fwd_d[lo:hi], fwd_i[lo:hi] = md[: hi - lo], mi[: hi - lo]
    # column-wise: for column block b, the parts are ranks b, b + C, ..., b + (R - 1) * C
    g, size = gather(rd, ri)
    rev_d = torch.empty((cols.n, k_rev), dtype=torch.float64, device=dev)
    rev_i = torch.empty((cols.n, k_rev), dtype=torch.int64, device=dev)
    for b in range(n_c):
        lo, hi = min(cols.n, b * per_c), min(cols.n, (b + 1) * per_c)
        if hi <= lo:
            continue
        base = b * size
        md, mi = merge(g[base:].view(torch.float64), g[base + per_c * k_rev:], n_r, n_c * size, k_rev,
                       per_c)
        rev_d[lo:hi], rev_i[lo:hi] = md[: hi - lo], mi[: hi - lo]
    return (fwd_d, fwd_i), (rev_d, rev_i)
