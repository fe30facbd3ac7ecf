"""``hubness_score`` on device -- kiez/analysis/estimation.py:197-351.

The k-occurrence histogram (integer, exact), the moment reductions, the hub /
antihub id lists and the Gini numerator are computed by the kernels in
csrc/analysis.cu; only the closing scalar formulas run on the host.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional, Union

import numpy as np
import torch

VALID_HUBNESS_MEASURES = [
    "all", "all_but_gini", "k_skewness", "k_skewness_truncnorm", "atkinson", "gini",
    "robinhood", "antihubs", "antihub_occurrence", "hubs", "hub_occurrence", "groupie_ratio",
    "k_occurrence",
]


def _truncnorm_third_moment(mean: float, std: float) -> float:
    """stats.truncnorm(a, b).moment(3), a=(0-mean)/std, b=(int64max-mean)/std
    (estimation.py:37-58) in closed form: scipy's recurrence
    m_k = pdf(a) a^(k-1) - pdf(b) b^(k-1) + (k-1) m_(k-2) with pdf(b) = 0 at that bound."""
    if not std > 0.0:
        return float("nan")
    a = (0.0 - mean) / std
    sf = 0.5 * math.erfc(a / math.sqrt(2.0))
    pa = math.exp(-0.5 * a * a) / math.sqrt(2.0 * math.pi) / sf
    return pa * a * a + 2.0 * pa


def hubness_score(nn_ind, target_samples: int, *, k: Optional[int] = None,
                  hub_size: float = 2.0, verbose: int = 0, return_value: str = "all_but_gini",
                  store_k_occurrence: bool = False) -> Union[float, dict]:
    """Hubness measures of a neighbour-id matrix; arguments as in the reference."""
    from . import _lib as lib

    was_numpy = isinstance(nn_ind, np.ndarray)
    if was_numpy:
        if not np.issubdtype(nn_ind.dtype, np.integer):
            if not np.all(np.isfinite(nn_ind)):
                raise ValueError("nn_ind must hold integer ids: no negative infinity / NaN allowed")
            nn_ind = nn_ind.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(nn_ind))
    else:
        t = nn_ind
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ind = t.to(device=dev, dtype=torch.int64).contiguous()
    n_rows, width = ind.shape
    n_test = target_samples
    if k is None:
        k = width
    elif k > width:
        k = width
        warnings.warn(f"k > nn_ind.shape[1], k will be set to {k}", stacklevel=2)
    st = None
    with torch.cuda.device(dev):
        st = lib.stream_ptr()
        rng = torch.empty(2, dtype=torch.int64, device=dev)
        lib.call("kb2_index_range", lib.ptr(ind), n_rows, ind.stride(0), k, lib.ptr(rng), st)
        _, max_id = rng.cpu().tolist()
        nbins = max(n_rows, int(max_id) + 1, 1)   # np.bincount(..., minlength=n_rows)
        hist = torch.empty(nbins, dtype=torch.int64, device=dev)
        lib.call("kb2_k_occurrence", lib.ptr(ind), n_rows, ind.stride(0), k, nbins, lib.ptr(hist), st)
        mom = torch.empty(10, dtype=torch.float64, device=dev)
        lib.call("kb2_hub_moments", lib.ptr(hist), nbins, 0.0, hub_size * k, lib.ptr(mom), st)
        total = mom[0].item()
        mean = total / nbins
        lib.call("kb2_hub_moments", lib.ptr(hist), nbins, mean, hub_size * k, lib.ptr(mom), st)
        (s1, s_e2, s_e3, s_abs, s_sqrt, mx, n_zero, n_hub, hub_sum, _s2) = mom.cpu().tolist()
        m2, m3 = s_e2 / nbins, s_e3 / nbins
        k_skewness = m3 / m2 ** 1.5 if m2 > 0 else float("nan")
        std1 = math.sqrt(s_e2 / (nbins - 1)) if nbins > 1 else float("nan")
        res = {
            "k_skewness": k_skewness,
            "k_skewness_truncnorm": _truncnorm_third_moment(mean, std1),
            "atkinson": float(1.0 - 1.0 / mean * (s_sqrt / nbins) ** 2) if mean > 0 else float("nan"),
            "gini": float("nan"),
            "robinhood": 0.5 * s_abs / s1 if s1 > 0 else float("nan"),
        }
        scratch = torch.empty((nbins + 1023) // 1024 + 2, dtype=torch.int64, device=dev)
        cnt = torch.empty(1, dtype=torch.int64, device=dev)

        def compact(mode, thresh, count):
            out = torch.empty(int(count), dtype=torch.int64, device=dev)
            lib.call("kb2_compact_ids", lib.ptr(hist), nbins, mode, float(thresh), lib.ptr(scratch),
                     lib.ptr(out), lib.ptr(cnt), st)
            return out

        antihubs = compact(0, 0.0, n_zero)
        hubs = compact(1, hub_size * k, n_hub)
        if return_value in ("gini", "all"):
            max_value = int(mx)
            gs = torch.empty(max_value + 4, dtype=torch.int64, device=dev)
            num = torch.empty(1, dtype=torch.int64, device=dev)
            lib.call("kb2_gini_numerator", lib.ptr(hist), nbins, max_value, lib.ptr(gs),
                     lib.ptr(num), st)
            res["gini"] = num.item() / (2 * nbins * s1) if s1 > 0 else float("nan")
    conv = (lambda x: x.cpu().numpy()) if was_numpy else (lambda x: x)
    res.update({
        "antihubs": conv(antihubs),
        "antihub_occurrence": n_zero / nbins,
        "hubs": conv(hubs),
        "hub_occurrence": hub_sum / k / n_test,
        "groupie_ratio": mx / n_test / k,
    })
    if store_k_occurrence:
        res["k_occurrence"] = conv(hist)
    if return_value == "all":
        return res
    if return_value == "all_but_gini":
        del res["gini"]
        return res
    return res[return_value]
