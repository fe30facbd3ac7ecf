"""Hubness reduction on device.

Same classes, constructor arguments, attribute names and error behaviour as
kiez/hubness_reduction/{base,csls,local_scaling,mutual_proximity,dis_sim}.py;
the arithmetic runs in the fused float64 CUDA kernels of csrc/rescale.cu
(``kb2_rescale_topk`` & co.).  ``transform`` keeps the reference contract
(unsorted (n, c) output); ``kneighbors`` uses the fused rescale+top-k call.

The numeric target is the reference's **numpy** branch (the one SklearnNN
drives): e.g. MutualProximity uses population std (ddof=0) and the survival
function -- not the torch branch's ddof=1 / fp32 ``1-cdf``
(mutual_proximity.py:98-103,170-182).
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np

from .neighbors import NNAlgorithm, check_is_fitted

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _lib():
    from . import _lib as lib

    return lib


def _as_device(x, device, dtype):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=dtype).contiguous()


def _device_of(algo):
    dev = getattr(algo, "device", None)
    return dev if dev is not None else torch.device("cuda", torch.cuda.current_device())


def _to_host(*tensors):
    """Device results -> numpy.  Large results go through page-locked buffers from torch's
    caching host allocator (one DMA each, no staging through the driver's bounce buffer; the
    first call pays the cudaHostAlloc, later calls reuse the cached blocks) and the returned
    arrays are views of those buffers."""
    if sum(t.numel() * t.element_size() for t in tensors) < (8 << 20):
        return tuple(t.cpu().numpy() for t in tensors)
    outs = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    torch.cuda.current_stream(tensors[0].device).synchronize()
    return tuple(h.numpy() for h in outs)


def _rows_on_device(algo, data, cache):
    """fp32 rows of an embedding matrix on the device (re-uses the backend's upload)."""
    if hasattr(algo, "_prepare"):
        return algo._prepare(data, cache=cache).raw
    return _as_device(data, _device_of(algo), torch.float32)



class HubnessReduction(ABC):
    """hubness_reduction/base.py:17-105."""

    def __init__(self, nn_algo: NNAlgorithm, verbose: int = 0, **kwargs):
        self.nn_algo = nn_algo
        self.verbose = verbose
        self._use_torch = False
        if nn_algo.n_candidates == 1:
            raise ValueError(
                "Cannot perform hubness reduction with a single candidate per query!"
            )

    @abstractmethod
    def _fit(self, neigh_dist, neigh_ind, source, target):
        ...

    @abstractmethod
    def transform(self, neigh_dist, neigh_ind, query):
        ...

    # fused rescale + top-k; subclasses override, default = transform then sort
    def _transform_topk(self, neigh_dist, neigh_ind, query, k):
        d, i = self.transform(neigh_dist, neigh_ind, query)
        return HubnessReduction._sort(d, i, k)

    def fit(self, source, target=None):
        self.nn_algo.fit(source, target)
        if target is None:
            target = source
        # REVERSE pass: every target row against the source index.  `query=target` is
        # passed explicitly, so in single-source mode self is NOT excluded (base.py:37-42)
        neigh_dist_t_to_s, neigh_ind_t_to_s = self.nn_algo.kneighbors(
            k=self.nn_algo.n_candidates, query=target, s_to_t=False, return_distance=True)
        if torch is not None and isinstance(neigh_dist_t_to_s, torch.Tensor):
            self._use_torch = True
        self._fit(neigh_dist_t_to_s, neigh_ind_t_to_s, source, target)

    def _set_k_if_needed(self, k: Optional[int] = None) -> int:
        c = self.nn_algo.n_candidates
        if k is None:
            warnings.warn(f"No k supplied, setting to n_candidates = {c}", stacklevel=2)
            return c
        if k > c:
            warnings.warn(f"k > n_candidates supplied! Setting to n_candidates = {c}",
                          stacklevel=2)
            return c
        return k

    @staticmethod
    def _sort(hubness_reduced_query_dist, query_ind, n_neighbors: int):
        """Top-k of the rescaled candidates, ascending (base.py:72-87) -> kb2_topk_rows.
        Like the reference, the result mirrors the input container: numpy in -> numpy out,
        torch in -> torch out on the input's device (CUDA tensors stay on the device, which is
        what the pipeline passes)."""
        lib = _lib()
        dist, ind = hubness_reduced_query_dist, query_ind
        was_numpy = isinstance(dist, np.ndarray)
        back_to = dist.device if torch.is_tensor(dist) and not dist.is_cuda else None
        dev = dist.device if torch.is_tensor(dist) and dist.is_cuda else torch.device(
            "cuda", torch.cuda.current_device())
        d = _as_device(dist, dev, torch.float64)
        i = _as_device(ind, dev, torch.int64)
        n, c = d.shape
        k = min(n_neighbors, c)
        with torch.cuda.device(dev):
            od = torch.empty((n, k), dtype=torch.float64, device=dev)
            oi = torch.empty((n, k), dtype=torch.int64, device=dev)
            if n:
                lib.call("kb2_topk_rows", lib.ptr(d), lib.ptr(i), n, c, 1, 0, k, lib.ptr(od),
                         lib.ptr(oi), lib.stream_ptr())
        if was_numpy:
            return od.cpu().numpy(), oi.cpu().numpy()
        if back_to is not None:
            return od.to(back_to), oi.to(back_to)
        return od, oi

    def _finish(self, dist, ind):
        """Mirror the caller's container type: numpy in -> numpy out (like Faiss+numpy)."""
        if isinstance(dist, np.ndarray):          # a user-defined transform that stayed on the host
            return dist, ind
        if getattr(self.nn_algo, "_input_is_numpy", False):
            # distributed runs replicate the result on every rank's device; with
            # `algorithm.host_result = "rank0"` only rank 0 pays the device-to-host copy (the
            # other ranks keep their device tensors)
            if getattr(self.nn_algo, "host_result", "all") == "rank0" and \
                    getattr(self.nn_algo, "distributed", False) and \
                    torch.distributed.get_rank() != 0:
                return dist, ind
            return _to_host(dist, ind)
        return dist, ind

    def kneighbors(self, k: Optional[int] = None):
        n_neighbors = self._set_k_if_needed(k)
        # FORWARD pass: query=None => self excluded in single-source mode (base.py:92-94)
        query_dist, query_ind = self.nn_algo.kneighbors(
            query=None, k=self.nn_algo.n_candidates, return_distance=True)
        d, i = self._transform_topk(query_dist, query_ind, self.nn_algo.source_, n_neighbors)
        return self._finish(d, i)


class NoHubnessReduction(HubnessReduction):
    """hubness_reduction/base.py:108-122: forward search only, no reverse pass."""

    def _fit(self, neigh_dist, neigh_ind, source, target):
        pass

    def fit(self, source, target=None):
        self.nn_algo.fit(source, target, only_fit_target=True)

    def transform(self, neigh_dist, neigh_ind, query):
        return neigh_dist, neigh_ind

    def kneighbors(self, k: Optional[int] = None):
        n_neighbors = self._set_k_if_needed(k)
        d, i = self.nn_algo.kneighbors(query=None, k=n_neighbors, return_distance=True)
        return self._finish(d, i)


class _DeviceRescale(HubnessReduction):
    """Shared plumbing of the four gather-type rescalers (one kb2_rescale_topk call)."""

    _mode: int = -1

    def _row_stats(self, dist, want_mean=False, want_sd=False, want_last=False, mp=False):
        lib = _lib()
        dev = dist.device
        n, c = dist.shape
        with torch.cuda.device(dev):
            mean = torch.empty(n, dtype=torch.float64, device=dev) if want_mean else None
            sd = torch.empty(n, dtype=torch.float64, device=dev) if want_sd else None
            last = torch.empty(n, dtype=torch.float64, device=dev) if want_last else None
            lib.call("kb2_row_stats", lib.ptr(dist), n, c, lib.ptr(mean), lib.ptr(sd),
                     lib.ptr(last), lib.stream_ptr())
        return mean, sd, last

    def _stats(self):
        raise NotImplementedError

    def _rescale(self, neigh_dist, neigh_ind, k):
        lib = _lib()
        stat_a, stat_b = self._stats()
        dev = stat_a.device
        d = _as_device(neigh_dist, dev, torch.float64)
        i = _as_device(neigh_ind, dev, torch.int64)
        n, c = d.shape
        width = c if k == 0 else k
        with torch.cuda.device(dev):
            od = torch.empty((n, width), dtype=torch.float64, device=dev)
            oi = torch.empty((n, width), dtype=torch.int64, device=dev)
            lib.call("kb2_rescale_topk", self._mode, lib.ptr(d), lib.ptr(i), n, c,
                     lib.ptr(stat_a), lib.ptr(stat_b), stat_a.shape[0], k, lib.ptr(od),
                     lib.ptr(oi), lib.stream_ptr())
        return od, oi

    def transform(self, neigh_dist, neigh_ind, query=None):
        self._check_fitted()
        return self._rescale(neigh_dist, neigh_ind, 0)

    def _transform_topk(self, neigh_dist, neigh_ind, query, k):
        self._check_fitted()
        return self._rescale(neigh_dist, neigh_ind, min(k, neigh_dist.shape[1]))


class CSLS(_DeviceRescale):
    """Cross-domain similarity local scaling (csls.py): 2 d - mean(d_row) - r_train[ind]."""

    _mode = 0

    def __repr__(self):
        return f"{self.__class__.__name__}(verbose = {self.verbose})"

    def _fit(self, neigh_dist, neigh_ind, source=None, target=None):
        dev = _device_of(self.nn_algo)
        self.r_dist_train_ = _as_device(neigh_dist, dev, torch.float64)
        self.r_ind_train_ = _as_device(neigh_ind, dev, torch.int64)
        # csls.py:90 recomputes this mean on every transform; it only depends on fit data
        self._r_train_mean, _, _ = self._row_stats(self.r_dist_train_, want_mean=True)
        return self

    def _check_fitted(self):
        check_is_fitted(self, "r_dist_train_")

    def _stats(self):
        return self._r_train_mean, None


class LocalScaling(_DeviceRescale):
    """Local scaling / NICDM (local_scaling.py)."""

    def __init__(self, method: str = "standard", **kwargs):
        super().__init__(**kwargs)
        self.method = method.lower()
        if self.method not in ["ls", "standard", "nicdm"]:
            raise ValueError(f"Internal: Invalid method {self.method}. Try 'ls' or 'nicdm'.")
        self._mode = 2 if self.method == "nicdm" else 1

    def __repr__(self):
        return f"{self.__class__.__name__}(method = {self.method}, verbose = {self.verbose})"

    def _fit(self, neigh_dist, neigh_ind, source, target):
        dev = _device_of(self.nn_algo)
        self.r_dist_t_to_s_ = _as_device(neigh_dist, dev, torch.float64)
        self.r_ind_t_to_s_ = _as_device(neigh_ind, dev, torch.int64)
        if self.method == "nicdm":
            self._r_stat, _, _ = self._row_stats(self.r_dist_t_to_s_, want_mean=True)
        else:
            _, _, self._r_stat = self._row_stats(self.r_dist_t_to_s_, want_last=True)
        return self

    def _check_fitted(self):
        check_is_fitted(self, "r_dist_t_to_s_")

    def _stats(self):
        return self._r_stat, None


class MutualProximity(_DeviceRescale):
    """Mutual proximity, Gaussian ('normal') or empiric (mutual_proximity.py)."""

    _mode = 3

    def __init__(self, method: str = "normal", **kwargs):
        super().__init__(**kwargs)
        if method not in ["exact", "empiric", "normal", "gaussi"]:
            raise ValueError(
                f'Mutual proximity method "{method}" not recognized. Try "normal" or "empiric".'
            )
        self.method = "empiric" if method in ["exact", "empiric"] else "normal"

    def __repr__(self):
        return f"{self.__class__.__name__}(method = {self.method}, verbose = {self.verbose})"

    def _fit(self, neigh_dist, neigh_ind, source, target):
        dev = _device_of(self.nn_algo)
        d = _as_device(neigh_dist, dev, torch.float64)
        self.n_train = d.shape[0]
        if self.method == "empiric":
            self.neigh_dist_t_to_s_ = d
            self.neigh_ind_t_to_s_ = _as_device(neigh_ind, dev, torch.int64)
        else:
            # numpy branch: nanmean / nanstd with ddof=0 (mutual_proximity.py:101-103)
            self.mu_t_to_s_, self.sd_t_to_s_, _ = self._row_stats(d, want_mean=True, want_sd=True)
        return self

    def _check_fitted(self):
        check_is_fitted(
            self, ["mu_t_to_s_", "sd_t_to_s_", "neigh_dist_t_to_s_", "neigh_ind_t_to_s_"],
            all_or_any=any)

    def _stats(self):
        return self.mu_t_to_s_, self.sd_t_to_s_

    def _rescale(self, neigh_dist, neigh_ind, k):
        if self.method == "normal":
            return super()._rescale(neigh_dist, neigh_ind, k)
        lib = _lib()
        rd, ri = self.neigh_dist_t_to_s_, self.neigh_ind_t_to_s_
        dev = rd.device
        d = _as_device(neigh_dist, dev, torch.float64)
        i = _as_device(neigh_ind, dev, torch.int64)
        n, c = d.shape
        width = c if k == 0 else k
        with torch.cuda.device(dev):
            od = torch.empty((n, width), dtype=torch.float64, device=dev)
            oi = torch.empty((n, width), dtype=torch.int64, device=dev)
            lib.call("kb2_mp_empiric_topk", lib.ptr(d), lib.ptr(i), n, c, lib.ptr(rd),
                     lib.ptr(ri), rd.shape[0], rd.shape[1], k, lib.ptr(od), lib.ptr(oi),
                     lib.stream_ptr())
        return od, oi


class DisSimLocal(HubnessReduction):
    """DisSimLocal (dis_sim.py): ||q-t||^2 - ||q-c_q||^2 - ||t-c_t||^2 with local centroids."""

    def __init__(self, squared: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.squared = squared
        metric = self.nn_algo.metric
        if metric in ["euclidean", "minkowski"]:
            self.squared = False
            if hasattr(self.nn_algo, "p") and self.nn_algo.p != 2:
                raise ValueError(
                    "DisSimLocal only supports squared Euclidean distances. If"
                    " the provided NNAlgorithm has a `p` parameter it must be"
                    f" set to p=2. Now it is p={self.nn_algo.p}"
                )
        elif metric in ["sqeuclidean"]:
            self.squared = True
        else:
            raise ValueError(
                f"DisSimLocal only supports squared Euclidean distances, not metric={metric}."
            )

    def __repr__(self):
        return f"{self.__class__.__name__}(squared = {self.squared})"

    def _fit(self, neigh_dist, neigh_ind, source, target):
        lib = _lib()
        algo = self.nn_algo
        dev = _device_of(algo)
        src = _rows_on_device(algo, source, True)
        tgt = _rows_on_device(algo, target, True)
        if src.dtype != tgt.dtype:
            src, tgt = src.to(torch.float64), tgt.to(torch.float64)
        ri = _as_device(neigh_ind, dev, torch.int64)
        m, c_rev = ri.shape
        d = src.shape[1]
        # multi-GPU: every rank fits a shard of the target rows (the centroid gathers are the
        # heavy part: m * c * d * 4 bytes) and the per-target scalars are all-gathered; the
        # centroids themselves stay sharded (`target_centroids_` = this rank's rows)
        lo, hi = 0, m
        if self._sharded():
            from .distributed import shard_bounds

            lo, hi = shard_bounds(m, torch.distributed.get_world_size(), torch.distributed.get_rank())
        with torch.cuda.device(dev):
            cent = torch.empty((hi - lo, d), dtype=torch.float64, device=dev)
            d2c = torch.empty(hi - lo, dtype=torch.float64, device=dev)
            if hi > lo:
                lib.call("kb2_dsl_fit", lib.ptr(src), src.shape[0], src.stride(0), lib.ptr(tgt[lo:hi]),
                         hi - lo, tgt.stride(0), d, src.element_size(), lib.ptr(ri[lo:hi]), c_rev,
                         lib.ptr(cent), lib.ptr(d2c), lib.stream_ptr())
        if self._sharded():
            from .distributed import all_gather_vector

            d2c = all_gather_vector(d2c, m)
        self.source_ = source
        self.target_ = target
        self._target_dev = tgt
        self.target_centroids_ = cent
        self.target_dist_to_centroids_ = d2c
        return self

    def _sharded(self) -> bool:
        """Row-sharded rescale: a distributed backend replicates the kNN results on every rank."""
        return bool(getattr(self.nn_algo, "distributed", False)) and \
            torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1

    def _raw(self, neigh_ind, query, lo=0, hi=None):
        lib = _lib()
        algo = self.nn_algo
        dev = _device_of(algo)
        q = _rows_on_device(algo, query, False)
        tgt = self._target_dev
        if q.dtype != tgt.dtype:
            q = q.to(tgt.dtype)
        i = _as_device(neigh_ind, dev, torch.int64)
        if hi is not None:
            q, i = q[lo:hi], i[lo:hi].contiguous()
        n, c = i.shape
        with torch.cuda.device(dev):
            raw = torch.empty((n, c), dtype=torch.float64, device=dev)
            gmin = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
            if n:
                lib.call("kb2_dsl_transform", lib.ptr(q), n, q.stride(0), lib.ptr(tgt), tgt.shape[0],
                         tgt.stride(0), q.shape[1], q.element_size(), lib.ptr(i), c,
                         lib.ptr(self.target_dist_to_centroids_), lib.ptr(raw), lib.ptr(gmin),
                         lib.stream_ptr())
        return raw, i, gmin

    def _run(self, neigh_ind, query, k):
        """Both stages; with a distributed backend the query rows are sharded and the global
        minimum (dis_sim.py:171-173) is agreed with one all-reduce(MIN)."""
        if not self._sharded():
            raw, i, gmin = self._raw(neigh_ind, query)
            return self._finish_topk(raw, i, gmin, k)
        from .distributed import dsl_transform_sharded

        return dsl_transform_sharded(
            neigh_ind.shape[0], lambda lo, hi: self._raw(neigh_ind, query, lo, hi),
            lambda raw, i, gmin: self._finish_topk(raw, i, gmin, k))

    def _finish_topk(self, raw, i, gmin, k):
        lib = _lib()
        dev = raw.device
        n, c = raw.shape
        width = c if k == 0 else k
        with torch.cuda.device(dev):
            od = torch.empty((n, width), dtype=torch.float64, device=dev)
            oi = torch.empty((n, width), dtype=torch.int64, device=dev)
            if n:
                lib.call("kb2_dsl_finish_topk", lib.ptr(raw), lib.ptr(i), n, c, lib.ptr(gmin),
                         int(bool(self.squared)), k, lib.ptr(od), lib.ptr(oi), lib.stream_ptr())
        return od, oi

    def transform(self, neigh_dist, neigh_ind, query):
        check_is_fitted(self, ["target_", "target_centroids_", "target_dist_to_centroids_"])
        return self._run(neigh_ind, query, 0)          # neigh_dist is ignored (dis_sim.py:152-157)

    def _transform_topk(self, neigh_dist, neigh_ind, query, k):
        check_is_fitted(self, ["target_", "target_centroids_", "target_dist_to_centroids_"])
        return self._run(neigh_ind, query, min(k, neigh_ind.shape[1]))
