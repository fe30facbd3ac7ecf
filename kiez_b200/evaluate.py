"""hits@k on device -- kiez/evaluate/eval_metrics.py:23-61."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Union

import numpy as np
import torch


def hits(nn_ind, gold: Union[Dict[int, int], "np.ndarray", "torch.Tensor"],
         k: Optional[Iterable[int]] = None) -> Dict[int, float]:
    """Share of gold pairs whose target id is among the first k neighbours of its source row.

    ``gold`` is either a mapping source row -> target id (the reference's form) or an array
    with gold[i] = target id of source row i (negative = not evaluated).  As in the reference
    the denominator is ``len(gold)`` for a mapping -- a key that is not a row of ``nn_ind``
    counts as a miss (eval_metrics.py:8-12,61) -- ``k=None`` means [1, 5, 10] and a list of
    rows or a dict row -> neighbour list is accepted for ``nn_ind`` (:57-60).
    """
    from . import _lib as lib

    ks = sorted(int(x) for x in ([1, 5, 10] if k is None else k))
    key_rows = None
    if isinstance(nn_ind, dict):                       # {source row: neighbour list}
        key_rows = list(nn_ind.keys())
        nn_ind = np.asarray([list(v) for v in nn_ind.values()], dtype=np.int64)
    elif isinstance(nn_ind, list):
        nn_ind = np.asarray(nn_ind, dtype=np.int64)
    dev = nn_ind.device if torch.is_tensor(nn_ind) and nn_ind.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    ind = (torch.from_numpy(np.ascontiguousarray(nn_ind)) if isinstance(nn_ind, np.ndarray)
           else nn_ind).to(device=dev, dtype=torch.int64)
    if ind.dim() != 2:
        raise ValueError(f"nn_ind must be 2-d (rows x neighbours), got shape {tuple(ind.shape)}")
    ind = ind.contiguous()
    n, width = ind.shape
    if isinstance(gold, dict):
        denominator = len(gold)
        g = np.full(n, -1, dtype=np.int64)
        row_of = {r: i for i, r in enumerate(key_rows)} if key_rows is not None else None
        for s_, t_ in gold.items():
            i = row_of.get(s_, -1) if row_of is not None else s_
            if isinstance(i, (int, np.integer)) and 0 <= i < n:      # other keys: counted, never hit
                g[int(i)] = int(t_)
        gold_t = torch.from_numpy(g)
    else:
        gold_t = torch.from_numpy(np.ascontiguousarray(gold)) if isinstance(gold, np.ndarray) \
            else torch.as_tensor(gold)
        gold_t = gold_t.reshape(-1).to(torch.int64)
        if gold_t.numel() < n:                         # rows without a gold entry are not evaluated
            pad = torch.full((n - gold_t.numel(),), -1, dtype=torch.int64, device=gold_t.device)
            gold_t = torch.cat([gold_t, pad])
        denominator = int((gold_t >= 0).sum().item())
        gold_t = gold_t[:n]
    gold_t = gold_t.to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        ks_t = torch.tensor(ks, dtype=torch.int32, device=dev)
        counts = torch.zeros(len(ks), dtype=torch.int64, device=dev)
        if n:
            lib.call("kb2_hits", lib.ptr(ind), n, ind.stride(0), width, lib.ptr(gold_t),
                     lib.ptr(ks_t), len(ks), lib.ptr(counts), lib.stream_ptr())
    counts = counts.cpu().tolist()
    return {kk: (cnt / denominator if denominator else 0.0) for kk, cnt in zip(ks, counts)}
