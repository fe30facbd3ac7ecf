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
