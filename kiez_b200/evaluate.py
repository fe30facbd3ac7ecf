"""hits@k on device -- kiez/evaluate/eval_metrics.py:23-61."""
from __future__ import annotations

from typing import Dict, Iterable, Union

import numpy as np
import torch


def hits(nn_ind, gold: Union[Dict[int, int], "np.ndarray", "torch.Tensor"],
         k: Iterable[int] = (1, 5, 10)) -> Dict[int, float]:
    """Share of evaluated source rows whose gold target id is among the first k neighbours.

    ``gold`` is either a mapping source row -> target id (the reference's dict form) or an
    array with gold[i] = target id of source row i (negative = not evaluated).
    """
    from . import _lib as lib

    dev = nn_ind.device if torch.is_tensor(nn_ind) and nn_ind.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    ind = (torch.from_numpy(np.ascontiguousarray(nn_ind)) if isinstance(nn_ind, np.ndarray)
           else nn_ind).to(device=dev, dtype=torch.int64).contiguous()
    n, width = ind.shape
    if isinstance(gold, dict):
        g = np.full(n, -1, dtype=np.int64)
        for s, t in gold.items():
            g[int(s)] = int(t)
        gold_t = torch.from_numpy(g)
    elif isinstance(gold, np.ndarray):
        gold_t = torch.from_numpy(np.ascontiguousarray(gold))
    else:
        gold_t = torch.as_tensor(gold)
    gold_t = gold_t.to(device=dev, dtype=torch.int64).contiguous()
    ks = [int(x) for x in k]
    with torch.cuda.device(dev):
        ks_t = torch.tensor(ks, dtype=torch.int32, device=dev)
        counts = torch.zeros(len(ks), dtype=torch.int64, device=dev)
        lib.call("kb2_hits", lib.ptr(ind), n, ind.stride(0), width, lib.ptr(gold_t),
                 lib.ptr(ks_t), len(ks), lib.ptr(counts), lib.stream_ptr())
    evaluated = int((gold_t >= 0).sum().item())
    counts = counts.cpu().tolist()
    return {kk: (cnt / evaluated if evaluated else 0.0) for kk, cnt in zip(ks, counts)}
