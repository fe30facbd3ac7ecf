"""Nearest-neighbour plugin API and the ``B200`` exact backend.

Mirrors, name for name, the reference's plugin interface
(kiez/neighbors/neighbor_algorithm_base.py:13-136): ``NNAlgorithm`` with
``fit`` / ``kneighbors`` and the abstract ``_fit`` / ``_kneighbors`` /
``valid_metrics``.  ``B200`` implements it with the CUDA library behind
include/kiez_b200.h: *prepare* (3xTF32 operand split, "index build"),
*candidate search* (tcgen05 tiles + fused top-c selection) and *exact finish*
(fp64 distances of the candidates, sorted) -- what
kiez/neighbors/exact/sklearn_nearest_neighbors.py:83-101 delegates to
scikit-learn's brute-force ``NearestNeighbors``.

All backend logic lives in ``B200Mixin`` so that ``kiez_b200.plugin`` can also
graft it onto the *real* ``kiez.neighbors.NNAlgorithm`` when kiez is installed.
"""
from __future__ import annotations

import os
import time
import warnings
from abc import ABC, abstractmethod
from typing import Any, Optional, Tuple

import numpy as np

try:  # torch is plumbing here: device memory, streams, torch.distributed
    import torch
except ImportError:  # pragma: no cover
    torch = None


class NotFittedError(ValueError, AttributeError):
    """Same bases as sklearn.exceptions.NotFittedError, which the reference raises."""


def check_is_fitted(obj, attributes, all_or_any=all):
    if not isinstance(attributes, (list, tuple)):
        attributes = [attributes]
    if not all_or_any([hasattr(obj, a) for a in attributes]):
        raise NotFittedError(
            f"This {type(obj).__name__} instance is not fitted yet. Call 'fit' with "
            "appropriate arguments before using this estimator."
        )


class NNAlgorithm(ABC):
    """Base class of nearest-neighbour backends (neighbor_algorithm_base.py:13)."""

    _ALLOWED_INPUT_TYPES: Tuple[Any, ...] = (np.ndarray,)

    def __init__(self, n_candidates, metric, n_jobs):
        self.n_candidates = n_candidates
        self.metric = metric
        self.n_jobs = n_jobs

    def _describe_source_target_fitted(self):
        if hasattr(self, "source_"):
            return (f" is fitted with: source.shape={self.source_.shape} and"
                    f" target.shape={self.target_.shape}")
        return " is unfitted"

    @property
    @abstractmethod
    def valid_metrics(self):
        ...

    @abstractmethod
    def _fit(self, data, is_source: bool) -> Any:
        ...

    @abstractmethod
    def _kneighbors(self, k, query, index, return_distance, is_self_querying):
        ...

    def _check_input_types(self, value):
        if not isinstance(value, tuple):
            value = (value,)
        allowed = self.__class__._ALLOWED_INPUT_TYPES
        if any(x is not None and not isinstance(x, allowed) for x in value):
            raise ValueError(
                f"Not implemented for input type(s) {[type(x) for x in value]}! "
                f"Only {allowed} allowed!"
            )

    def fit(self, source, target=None, only_fit_target: bool = False):
        """Index the data (neighbor_algorithm_base.py:53-96)."""
        self._check_input_types((source, target))
        self.source_equals_target = target is None
        if self.source_equals_target:
            self.source_index = self._fit(source, True)
            self.target_index = self.source_index
            target = source
        else:
            if source.shape[1] != target.shape[1]:
                raise ValueError(
                    "Expected source and target to have the same number of features,"
                    f" but got source.shape: {source.shape} and target.shape: {target.shape}"
                )
            if only_fit_target:
                self.target_index = self._fit(target, True)
            else:
                self.source_index = self._fit(source, True)
                self.target_index = self._fit(target, False)
        self.source_ = source
        self.target_ = target

    def _check_k_value(self, k: int, needed_space) -> int:
        if not np.issubdtype(type(k), np.integer):
            raise TypeError(f"k does not take {type(k)} value, enter integer value")
        if k <= 0:
            raise ValueError(f"Expected k > 0. Got {k}")
        if k > needed_space:
            warnings.warn(
                f"k={k} is larger than number of samples in indexed space.\n"
                f"Setting to k={needed_space}",
                stacklevel=2,
            )
            return needed_space
        return k

    def kneighbors(self, k=None, query=None, s_to_t=True, return_distance=True):
        """neighbor_algorithm_base.py:116-136."""
        check_is_fitted(self, ["source_index", "target_index"], all_or_any=any)
        k = self.n_candidates if k is None else k
        is_self_querying = query is None and self.source_equals_target
        if s_to_t:
            query = self.source_ if query is None else query
            index = self.target_index
            needed_space = self.target_.shape[0]
        else:
            query = self.target_ if query is None else query
            index = self.source_index
            needed_space = self.source_.shape[0]
        k = self._check_k_value(k, needed_space)
        return self._kneighbors(k=k, query=query, index=index,
                                return_distance=return_distance,
                                is_self_querying=is_self_querying)


# ---------------------------------------------------------------------------
# the B200 backend
# ---------------------------------------------------------------------------

_METRIC_CODES = {
    "euclidean": 0, "l2": 0, "minkowski": 0,
    "sqeuclidean": 1,
    "cosine": 2,
}


class _PendingRows:
    """Bookkeeping of a matrix whose rows are still being uploaded (kiez_b200/upload.py): the
    first `done` chunks are on the device AND prepared."""

    __slots__ = ("algo", "upload", "done")

    def __init__(self, algo, upload):
        self.algo, self.upload, self.done = algo, upload, 0


class PreparedRows:
    """Device-resident operands of one embedding matrix: the raw fp32 rows (exact
    finish), their 3xTF32 split and selection term (candidate search).

    A matrix that arrives from the host is prepared chunk by chunk as its upload proceeds:
    `ensure(upto)` makes the first `upto` rows usable on the current stream (all rows when
    omitted); `rows(lo, hi)` does so for the shard it returns, `take` / `keymax` / `errmax`
    for the whole matrix.  Every consumer that reads the full tensors calls `ensure()` first."""

    __slots__ = ("raw", "hi", "lo", "key", "sqnorm", "_keymax", "err", "_errmax", "n", "d", "dpad",
                 "base", "_owner", "_pending", "_root", "presample")

    def __init__(self, raw, hi, lo, key, sqnorm, base=0, owner=None, keymax=None, err=None,
                 errmax=None, root=None):
        self.raw, self.hi, self.lo, self.key, self.sqnorm = raw, hi, lo, key, sqnorm
        self._keymax = keymax   # device scalar >= max(key): input of the screen's completeness proof
        # TF32 rounding error of the rows: err [n] >= ||w - hi||^2, errmax (device scalar) its maximum
        self.err, self._errmax = err, errmax
        self.n, self.d = raw.shape
        self.dpad = hi.shape[1]
        self.base = base        # global id of row 0 (multi-GPU shards)
        self._owner = owner     # keeps the user's array alive while its id() is a cache key
        self._pending = None    # _PendingRows while the upload / preparation is in progress
        self._root = root       # the matrix this is a shard of (its maxima cover the shard)
        self.presample = None   # ((count, step), PreparedRows) of a strided row sample uploaded first

    def ensure(self, upto=None):
        """Rows [0, upto) (default: all) are uploaded and prepared once the work enqueued on the
        current stream so far has run."""
        if self._root is not None:
            # a shard: its parent prepares a prefix of chunks covering it
            self._root.ensure(None if upto is None else self._root_offset() + upto)
            return self
        p = self._pending
        if p is None:
            return self
        up = p.upload
        last = len(up.bounds) - 1 if upto is None else up.chunk_of(max(0, min(self.n, upto) - 1))
        stream = torch.cuda.current_stream(self.raw.device)
        while p.done <= last:
            up.wait(p.done, stream)
            lo, hi = up.bounds[p.done]
            p.algo._prepare_range(self, lo, hi)
            p.done += 1
        if p.done == len(up.bounds):
            self._pending = None
            p.algo._finish_maxima(self)
        return self

    def _root_offset(self):
        return self.base - self._root.base

    @property
    def keymax(self):
        root = self._root if self._root is not None else self
        root.ensure()
        return root._keymax

    @property
    def errmax(self):
        root = self._root if self._root is not None else self
        root.ensure()
        return root._errmax

    def take(self, idx):
        """The rows `idx` (a device int64 tensor) gathered into a new contiguous set."""
        self.ensure()
        return PreparedRows(self.raw.index_select(0, idx), self.hi.index_select(0, idx),
                            self.lo.index_select(0, idx), self.key.index_select(0, idx),
                            None if self.sqnorm is None else self.sqnorm.index_select(0, idx),
                            base=0, owner=self._owner, keymax=self.keymax,
                            err=None if self.err is None else self.err.index_select(0, idx),
                            errmax=self.errmax)

    def rows(self, lo, hi):
        """A contiguous row shard (views, no copy)."""
        self.ensure(hi)
        root = self._root if self._root is not None else self
        return PreparedRows(self.raw[lo:hi], self.hi[lo:hi], self.lo[lo:hi], self.key[lo:hi],
                            None if self.sqnorm is None else self.sqnorm[lo:hi],
                            base=self.base + lo, owner=self._owner,
                            err=None if self.err is None else self.err[lo:hi], root=root)


def candidate_capacity(c: int) -> int:
    """Length of the per-row candidate list the search kernel keeps for `c` wanted
    neighbours: a margin absorbs rank swaps between the fp32 3xTF32 selection key
    and the exact fp64 distance at the c-th / (c+1)-th boundary."""
    return min(128, ((c + max(6, c // 8)) + 7) // 8 * 8)


class B200Mixin:
    """Implementation of the exact B200 backend; combine with an NNAlgorithm base."""

    valid_metrics = tuple(_METRIC_CODES)

    def __init__(self, n_candidates: int = 5, metric: str = "euclidean", p: int = 2,
                 device: Optional[Any] = None, impl: str = "auto", center: bool = True,
                 distributed: Optional[bool] = None, fused: Any = "auto",
                 precision: str = "auto", n_jobs=None, shard_mode: Optional[str] = None):
        if torch is None or not torch.cuda.is_available():
            raise ImportError(
                "The B200 backend needs PyTorch with a CUDA device (sm_100a); there is no "
                "CPU fallback."
            )
        from . import _lib  # raises ImportError when the CUDA library is not built

        if metric not in self.__class__.valid_metrics:
            raise ValueError(f"Unknown metric {metric}, please use one of {self.valid_metrics}")
        if metric == "minkowski" and p != 2:
            raise ValueError("B200 is an exact contraction backend: minkowski needs p=2")
        if impl not in ("auto", "tc", "tc1", "simt"):
            raise ValueError(f"impl must be 'auto', 'tc', 'tc1' or 'simt', got {impl!r}")
        if precision not in ("auto", "tf32x3", "screen"):
            raise ValueError(f"precision must be 'auto', 'tf32x3' or 'screen', got {precision!r}")
        if isinstance(n_candidates, (int, np.integer)) and n_candidates > _lib.lib.kb2_max_candidates():
            raise ValueError(
                f"B200 keeps at most {_lib.lib.kb2_max_candidates()} candidates per query "
                f"(the lists live in shared memory), got n_candidates={n_candidates}")
        super().__init__(n_candidates=n_candidates, metric=metric, n_jobs=n_jobs)
        self.p = p
        self.impl = impl
        self.precision = precision
        self.search_stats = {"screen_rows": 0, "screen_unverified": 0}
        self.center = center
        self._lib = _lib
        self._metric_code = _METRIC_CODES[metric]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1
        self.distributed = bool(distributed)
        # numpy callers in distributed runs: "all" = every rank copies the (replicated) result
        # to its host, "rank0" = only rank 0 does (the others keep device tensors)
        self.host_result = "all"
        # distributed dual-direction pass: "rows" = every rank takes a shard of the source rows and
        # the ranks agree on the column thresholds (default), "cols" = column (target) shards
        if shard_mode is None:
            shard_mode = os.environ.get("KB2_SHARD_MODE", "rows")
        if shard_mode not in ("rows", "cols"):
            raise ValueError(f"shard_mode must be 'rows' or 'cols', got {shard_mode!r}")
        self.shard_mode = shard_mode
        if fused not in ("auto", True, False):
            raise ValueError(f"fused must be 'auto', True or False, got {fused!r}")
        self.fused = fused
        self._fused_forward = None     # forward result cached by the dual-direction pass
        self._screen_ok = None         # verdict of the screen probe for this fit (None: not probed)
        self._screen_boost = False     # the probe asked for longer candidate lists
        self._prepared = {}
        self._center_vec = None
        self._input_is_numpy = False

    def __repr__(self):
        return (f"{self.__class__.__name__}(n_candidates={self.n_candidates},"
                f"metric={self.metric},impl={self.impl},device={self.device},"
                f"distributed={self.distributed})")

    # -- NNAlgorithm hooks ---------------------------------------------------
    def fit(self, source, target=None, only_fit_target: bool = False):
        self._prepared = {}
        self._center_vec = None
        self._fused_forward = None
        self._screen_ok = None
        self._screen_boost = False
        self._input_is_numpy = isinstance(source, np.ndarray)
        self._t_fit_start = time.perf_counter()     # origin of the upload timeline (bench.py e2e)
        self._start_uploads(source, target, only_fit_target)
        return super().fit(source, target, only_fit_target=only_fit_target)

    def _fit(self, data, is_source: bool):
        return self._prepare(data, cache=True, lazy=True)

    # Host inputs (what kiez passes): upload on a background thread in row chunks, overlapped
    # with the searches (kiez_b200/upload.py).  Order: the strided row sample of the source that
    # seeds the column thresholds (small), then the target -- the threshold search consumes it
    # chunk by chunk as it arrives --, then the source, whose row segments the dual-direction
    # pass consumes in order.  KB2_ASYNC_UPLOAD=0 restores the plain synchronous copies.
    ASYNC_UPLOAD_MIN_BYTES = 8 << 20

    def _start_uploads(self, source, target, only_fit_target):
        from . import upload

        self._uploader = None
        self._pending_uploads = {}
        if os.environ.get("KB2_ASYNC_UPLOAD", "1") == "0":
            return
        if self.distributed:
            # every rank uploads its 1 / world slice of each host matrix (the all-gather over
            # NVLink completes it, distributed.upload_sharded): start both slices now, through
            # the pinned staging ring, so that the copies overlap each other and the first kernels
            from .distributed import shard_slice

            up = upload.Uploader()
            with torch.cuda.device(self.device):
                for m in ((source, target) if not only_fit_target else (target,)):
                    if m is None or not upload.eligible(m) or m.shape[0] < 4096 or \
                            ("slice", id(m)) in self._pending_uploads:
                        continue
                    lo, hi, _per = shard_slice(m.shape[0])
                    if hi > lo:
                        self._pending_uploads[("slice", id(m))] = up.add(
                            upload.HostUpload(m[lo:hi], self.device))
            if up.jobs:
                up.start()
                self._uploader = up
            return
        mats = [m for m in ((target, source) if not only_fit_target else (target,))
                if m is not None and upload.eligible(m)
                and m.shape[0] * m.shape[1] * 4 >= self.ASYNC_UPLOAD_MIN_BYTES]
        if not mats or (source is not None and target is not None
                        and tuple(source.shape[1:]) != tuple(target.shape[1:])):
            return
        cosine = self._metric_code == self._lib.METRIC_COSINE
        first = source if not only_fit_target else target      # the matrix that defines the centre
        # the strided row sample of the source goes first (it seeds the column thresholds); when
        # it exists the centre is its mean, taken ON THE DEVICE once it has arrived (_prepare):
        # a float64 mean over a strided view of a 1 GB host matrix costs ~25 ms of host time
        # during which nothing would be uploading yet
        with_sample = any(m is source for m in mats) and target is not None and not only_fit_target
        if self.center and not cosine and upload.eligible(first) and not with_sample:
            # any vector near the mean serves (distances are translation invariant): the mean
            # of a strided row sample, so that no pass over the whole host matrix is needed
            host = first if isinstance(first, np.ndarray) else first.numpy()
            stride = max(1, host.shape[0] // 65536)
            self._center_vec = torch.from_numpy(
                host[::stride].mean(axis=0, dtype=np.float64).astype(np.float32)).to(self.device)
        up = upload.Uploader()
        with torch.cuda.device(self.device):
            if with_sample:
                cap = candidate_capacity(self.n_candidates)
                n_s = self._fused_sample_rows(source.shape[0], cap)
                step = max(1, source.shape[0] // n_s)
                self._pending_uploads[("sample", id(source))] = (
                    (n_s, step), up.add(upload.HostUpload(source, self.device, rows=(n_s, step))))
            for m in mats:
                self._pending_uploads[id(m)] = up.add(upload.HostUpload(m, self.device))
        up.start()
        self._uploader = up

    def _kneighbors(self, k, query, index, return_distance, is_self_querying):
        q = self._prepare(query, cache=False, lazy=True)
        if is_self_querying and k > index.n - 1:
            # sklearn raises for n_neighbors > n_samples_fit - 1 with X=None
            raise ValueError(
                f"Expected n_neighbors <= n_samples_fit - 1, but n_neighbors = {k}, "
                f"n_samples_fit = {index.n}"
            )
        source_index = getattr(self, "source_index", None)
        target_index = getattr(self, "target_index", None)
        cached = self._fused_forward
        if (cached is not None and index is target_index and q is source_index
                and cached[0] == (k, bool(is_self_querying))):
            dist, ind = cached[1]                        # forward pass already done by the fused pass
        elif (index is source_index and q is target_index and not is_self_querying
              and self._use_fused(source_index, target_index, k)):
            # kiez's reverse pass (hubness_reduction/base.py:37-42): produce it together with the
            # forward pass that HubnessReduction.kneighbors will ask for next (base.py:92-94)
            single = bool(getattr(self, "source_equals_target", False))
            k_fwd = min(self.n_candidates, target_index.n - (1 if single else 0))
            if self.distributed:
                from .distributed import sharded_knn_both, sharded_knn_both_rows

                world = torch.distributed.get_world_size()
                cap = candidate_capacity(max(k, self.n_candidates))
                # rows sharded (default): needs a few full tiles of rows per rank; else columns
                by_rows = self.shard_mode == "rows" and source_index.n // world >= max(1024, 8 * cap)
                shard = sharded_knn_both_rows if by_rows else sharded_knn_both
                fwd, rev = shard(self, source_index, target_index, k_fwd, k, single)
            else:
                fwd, rev = self.search_both(source_index, target_index, k_fwd, k,
                                            exclude_self_rows=single)
            self._fused_forward = ((k_fwd, single), fwd)
            dist, ind = rev
        elif self.distributed:
            from .distributed import sharded_knn

            dist, ind = sharded_knn(self, q, index, k, is_self_querying)
        else:
            dist, ind = self.search(q, index, k, exclude_self=is_self_querying)
        return (dist, ind) if return_distance else ind

    # -- device pipeline -------------------------------------------------------
    def to_device(self, data):
        """C-contiguous rows on this backend's device: fp32 (the production dtype), or
        float64 when the caller passes float64 (the README example does), in which case
        the candidate search still runs on the fp32-rounded split but the exact finish
        reads the caller's float64 values.

        Distributed mode, host input: every rank holds the same host array (the contract of
        INTEGRATION.md section 4), so each rank uploads only its 1/world slice over PCIe and
        the slices are exchanged with one NCCL all-gather over NVLink -- world x less
        host-to-device traffic per rank than every rank uploading the whole matrix."""
        on_host = isinstance(data, np.ndarray) or not data.is_cuda
        if on_host and self.distributed and data.ndim == 2 and data.shape[0] >= 4096:
            return self._upload_sharded(data)
        if isinstance(data, np.ndarray):
            keep64 = data.dtype == np.float64
            t = torch.from_numpy(np.ascontiguousarray(
                data, dtype=np.float64 if keep64 else np.float32))
            return t.to(self.device, non_blocking=True)
        keep64 = data.dtype == torch.float64
        return data.to(device=self.device,
                       dtype=torch.float64 if keep64 else torch.float32).contiguous()

    def _upload_sharded(self, data):
        from .distributed import upload_sharded

        job = getattr(self, "_pending_uploads", {}).pop(("slice", id(data)), None)
        mine = None
        if job is not None:                      # this rank's slice is (being) uploaded already
            stream = torch.cuda.current_stream(self.device)
            for chunk in range(len(job.bounds)):
                job.wait(chunk, stream)
            mine = job.dev
        return upload_sharded(data, self.device, mine=mine)

    def _prepare(self, data, cache: bool, lazy: bool = False) -> PreparedRows:
        """The prepared operands of `data` (cached per fitted matrix).  `lazy`: a matrix whose
        upload is still in progress is returned as it is -- the caller `ensure`s the rows it
        reads; otherwise every row is usable on return."""
        hit = self._prepared.get(id(data))
        if hit is not None:
            return hit if lazy else hit.ensure()
        lib = self._lib
        pending = getattr(self, "_pending_uploads", {}).pop(id(data), None)
        raw = pending.dev if pending is not None else self.to_device(data)
        if raw.dim() != 2:
            raise ValueError(f"Expected a 2-d embedding matrix, got shape {tuple(raw.shape)}")
        n, d = raw.shape
        cosine = self._metric_code == lib.METRIC_COSINE
        sample = getattr(self, "_pending_uploads", {}).pop(("sample", id(data)), None) \
            if pending is not None else None
        if self._center_vec is None and self.center and not cosine and sample is not None:
            # centre = mean of the row sample that was uploaded first (see _start_uploads)
            job = sample[1]
            stream = torch.cuda.current_stream(self.device)
            for chunk in range(len(job.bounds)):
                job.wait(chunk, stream)
            self._center_vec = job.dev.mean(dim=0, dtype=torch.float64).to(torch.float32).contiguous()
        if self._center_vec is None and self.center and not cosine and n > 0:
            if pending is not None:
                raise RuntimeError("kiez_b200: the centre of a matrix that is still uploading must "
                                   "come from its host copy or its row sample (_start_uploads)")
            # distances are translation invariant; centring keeps ||y||^2 - 2 q.y well
            # conditioned in fp32 for embeddings far from the origin.  Any vector near the mean
            # serves: the mean of a strided row sample (65536 to 131071 rows of a large matrix)
            stride = max(1, n // 65536)
            self._center_vec = raw[::stride].mean(dim=0, dtype=torch.float64).to(torch.float32).contiguous()
        dpad = lib.lib.kb2_padded_dim(d)
        with torch.cuda.device(self.device):
            hi = torch.empty((n, dpad), dtype=torch.float32, device=self.device)
            lo = torch.empty((n, dpad), dtype=torch.float32, device=self.device)
            key = torch.empty((n,), dtype=torch.float32, device=self.device)
            sqn = torch.empty((n,), dtype=torch.float64, device=self.device) if cosine else None
            err = torch.empty((n,), dtype=torch.float32, device=self.device)
            keymax = torch.empty((1,), dtype=torch.float32, device=self.device)
            errmax = torch.empty((1,), dtype=torch.float32, device=self.device)
        prep = PreparedRows(raw, hi, lo, key, sqn, owner=data, keymax=keymax, err=err, errmax=errmax)
        if pending is not None:
            prep._pending = _PendingRows(self, pending)
            if sample is not None:
                prep.presample = (sample[0], self._prepare_upload(sample[1], owner=data))
            if not lazy:
                prep.ensure()
        else:
            if n:
                self._prepare_range(prep, 0, n)
                if cosine and raw.dtype == torch.float64:
                    prep.sqnorm = (raw * raw).sum(dim=1)      # exact norms of the float64 rows
            self._finish_maxima(prep)
        if cache:
            self._prepared[id(data)] = prep
        return prep

    def _prepare_upload(self, up, owner) -> PreparedRows:
        """PreparedRows over the device buffer of an upload in progress (prepared lazily)."""
        n, d = up.dev.shape
        dpad = self._lib.lib.kb2_padded_dim(d)
        cosine = self._metric_code == self._lib.METRIC_COSINE
        dev = self.device
        with torch.cuda.device(dev):
            prep = PreparedRows(up.dev, torch.empty((n, dpad), dtype=torch.float32, device=dev),
                                torch.empty((n, dpad), dtype=torch.float32, device=dev),
                                torch.empty((n,), dtype=torch.float32, device=dev),
                                torch.empty((n,), dtype=torch.float64, device=dev) if cosine else None,
                                owner=owner,
                                keymax=torch.empty((1,), dtype=torch.float32, device=dev),
                                err=torch.empty((n,), dtype=torch.float32, device=dev),
                                errmax=torch.empty((1,), dtype=torch.float32, device=dev))
        prep._pending = _PendingRows(self, up)
        return prep

    def _prepare_range(self, prep: PreparedRows, lo: int, hi: int) -> None:
        """kb2_prepare_rows + kb2_split_error_terms for rows [lo, hi) of `prep` (the index
        build, chunk by chunk while a host matrix is still arriving)."""
        lib = self._lib
        raw = prep.raw
        raw32 = raw[lo:hi] if raw.dtype == torch.float32 else raw[lo:hi].to(torch.float32)
        cosine = self._metric_code == lib.METRIC_COSINE
        with torch.cuda.device(self.device):
            scratch = torch.empty((1,), dtype=torch.float32, device=self.device)
            lib.call("kb2_prepare_rows", lib.ptr(raw32), hi - lo, prep.d, raw32.stride(0),
                     None if cosine else lib.ptr(self._center_vec), self._metric_code,
                     lib.ptr(prep.hi[lo:hi]), lib.ptr(prep.lo[lo:hi]), prep.dpad,
                     lib.ptr(prep.key[lo:hi]),
                     None if prep.sqnorm is None else lib.ptr(prep.sqnorm[lo:hi]), lib.stream_ptr())
            lib.call("kb2_split_error_terms", lib.ptr(prep.lo[lo:hi]), hi - lo, prep.dpad,
                     lib.ptr(prep.err[lo:hi]), lib.ptr(scratch), lib.stream_ptr())

    def _finish_maxima(self, prep: PreparedRows) -> None:
        """max ||w||^2 and max ||w - hi||^2 over all rows: inputs of the completeness proof."""
        lib = self._lib
        with torch.cuda.device(self.device):
            lib.call("kb2_max_f32", lib.ptr(prep.key), prep.n, lib.ptr(prep._keymax), lib.stream_ptr())
            lib.call("kb2_max_f32", lib.ptr(prep.err), prep.n, lib.ptr(prep._errmax), lib.stream_ptr())

    # -- screen (1xTF32 proposals + completeness proof) ------------------------------
    def _use_screen(self, q: PreparedRows, y: PreparedRows, cap: int, dual: bool) -> bool:
        """Whether the candidate search runs as the 1xTF32 screen (knn_screen.cu).  The proof in
        the exact finish assumes the operands are the caller's fp32 values (float64 callers keep
        the 3xTF32 search) and needs a full list per row to say anything."""
        if self.precision == "tf32x3" or self.impl not in ("auto", "tc"):
            return False
        if self.precision == "auto" and self._screen_ok is False:
            return False                 # the probe of this fit found the proof failing too often
        if q.raw.dtype != torch.float32 or y.raw.dtype != torch.float32:
            return False
        if q.dpad != y.dpad or y.n < 4 * cap or q.n < 1:
            return False
        return self._lib.lib.kb2_screen_stages(q.dpad, cap, int(dual)) > 0

    # The screen pays off while the proof holds for most rows.  On data whose neighbour gaps are
    # around the TF32 error bound (tight clusters of normalised vectors) many rows would be
    # searched twice, so with precision="auto" the first rows of a fit are a PROBE:
    #   <= SCREEN_BOOST_UNVERIFIED of them unproven -> carry on;
    #   more -> restart with LONGER lists (`_boosted_capacity`: the proof needs the gap between
    #           the k-th and the cap-th neighbour to exceed E, and that gap grows with cap);
    #   still > SCREEN_MAX_UNVERIFIED with the longer lists -> the rest of the fit uses 3xTF32.
    SCREEN_BOOST_UNVERIFIED = 0.04
    SCREEN_MAX_UNVERIFIED = 0.25
    SCREEN_PROBE_ROWS = 18944            # one-direction search: 2 x 128 rows for each of 74 CTA pairs

    @staticmethod
    def _boosted_capacity(cap: int) -> int:
        """Longer candidate lists for fits whose probe leaves too many rows unproven."""
        return min(64, (cap * 3 // 2 + 7) // 8 * 8) if cap < 64 else cap

    def _capacity(self, c: int, q: PreparedRows = None, dual: bool = False) -> int:
        """List length of the candidate search for `c` wanted neighbours (boosted by the probe)."""
        cap = candidate_capacity(c)
        if self._screen_boost and self.precision != "tf32x3":
            boosted = self._boosted_capacity(cap)
            if q is None or self._lib.lib.kb2_screen_stages(q.dpad, boosted, int(dual)) > 0:
                cap = boosted
        return cap

    def _screen_verdict(self, unverified, cap: int = 0, dpad: int = 0, dual: bool = False,
                        comm=None) -> bool:
        """Record (once per fit and list length) whether the screen carries on as configured,
        from the `unverified` flags of the probe rows; one host sync.  False = start over
        (`_screen_boost` or `_screen_ok` changed)."""
        if self.precision != "auto":
            return True
        if self._screen_ok is None:
            if unverified is None:
                frac = 0.0
            elif unverified.dim() == 0:                    # a fraction from _probe_fraction
                frac = float(unverified)
            else:
                frac = float(unverified.to(torch.float32).mean()) if unverified.numel() else 0.0
            if comm is not None:
                frac = comm.max_scalar(frac)       # every rank must take the same branch
            self.search_stats.setdefault("screen_probe_unverified", []).append(frac)
            can_boost = (not self._screen_boost and cap > 0 and self._boosted_capacity(cap) > cap
                         and self._lib.lib.kb2_screen_stages(dpad, self._boosted_capacity(cap),
                                                             int(dual)) > 0)
            if frac > self.SCREEN_BOOST_UNVERIFIED and can_boost:
                self._screen_boost = True          # verdict stays open: the longer lists are probed next
                return False
            self._screen_ok = frac <= self.SCREEN_MAX_UNVERIFIED
        return self._screen_ok

    def _probe_fraction(self, unverified, q: PreparedRows, y: PreparedRows, keys, lists: int,
                        cap: int, out_d, k: int):
        """Fraction of the probe rows a SINGLE list of `cap` entries would leave unproven (device
        scalar).  A probe over few query tiles searches the index in `lists` independent ranges;
        the proof then runs with tau = min over the lists' cap-th keys, about the (lists * cap)-th
        best key overall -- far easier to pass than the cap-th best that the chained launches
        of the rest of the fit (and the column side) prove against.  So the verdict is taken
        on what those would see: tau_1 = cap-th smallest key of the union of the lists."""
        if lists <= 1 or unverified.numel() == 0:
            return unverified.to(torch.float32).mean() if unverified.numel() else \
                torch.zeros((), device=self.device)
        lib = self._lib
        tau1 = torch.kthvalue(keys.view(q.n, lists * cap), cap, dim=1).values.double()
        up = 1.000001
        cosine = self._metric_code == lib.METRIC_COSINE
        one = torch.ones((), dtype=torch.float64, device=self.device)
        qn2 = one.expand(q.n) if cosine else q.key.double()
        ym2 = one if cosine else y.keymax.double()
        qn, ym = qn2.sqrt() * up, ym2.sqrt() * up
        dq, dym = q.err.double().sqrt() * up, y.errmax.double().sqrt() * up
        E = 2.0 * (dq * ym + qn * (1.0 + 2.0 ** -11) * dym + self._eps_acc(q.dpad) * qn * ym) \
            + 4.76837158203125e-07 * (ym2 + 2.0 * qn * ym)
        kd = out_d[:, k - 1]
        if cosine:
            ok = 2.0 * (kd - 1.0) + 1e-9 < tau1 - E - 1e-6
        else:
            d2 = kd * kd if self._metric_code == lib.METRIC_EUCLIDEAN else kd
            ok = d2 * (1.0 + 1e-12) < tau1 - E + qn2 * (1.0 - 2.4e-7)
        return 1.0 - (ok | torch.isinf(tau1)).to(torch.float32).mean()

    def _eps_acc(self, dpad: int) -> float:
        """Bound of the accumulation error of <hi(q), hi(y)> relative to |q| |y|: the TF32
        products are exact in fp32, dpad additions of unknown order (truncation allowed), plus
        the fp32 rounding of the centring.  The operand rounding itself enters the proof
        through the measured error norms (kb2_split_error_terms), see refine.cu."""
        return dpad * 2.0 ** -22 + 2.0 ** -21

    def _screen_plan(self, nq: int, ny: int, dpad: int, cap: int):
        import ctypes as C

        lib = self._lib
        steps, chained = C.c_int(0), C.c_int(0)
        sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        lib.call("kb2_screen_plan", nq, ny, dpad, cap, sm, C.addressof(steps), C.addressof(chained))
        lib.launch_counter -= 1          # host-only entry point
        return steps.value, chained.value

    def _screen_search(self, q: PreparedRows, y: PreparedRows, cap: int, dual=None):
        """Launch the screen: returns (cand [nq][L], keys [nq][L], lists per row)."""
        lib = self._lib
        dev = self.device
        steps, chained = self._screen_plan(q.n, y.n, q.dpad, cap)
        lists = 1 if chained else steps
        cand = torch.empty((q.n, lists * cap), dtype=torch.int32, device=dev)
        ckey = torch.empty((q.n, lists * cap), dtype=torch.float32, device=dev)
        flag = torch.empty(((q.n + 127) // 128 * 4,), dtype=torch.int32, device=dev) if chained else None
        prof = getattr(self, "_profile", None)       # bench.py: CUDA events around the searches
        if prof is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        if dual is None:
            lib.call("kb2_knn_screen", lib.ptr(q.hi), None, q.n, lib.ptr(y.hi), lib.ptr(y.key), y.n,
                     q.dpad, cap, steps, chained, lib.ptr(cand), lib.ptr(ckey), lib.ptr(flag),
                     None, None, None, 0, 0, lib.stream_ptr())
        else:
            tau, col_cnt, col_buf, col_cap, row_id_base = dual
            lib.call("kb2_knn_screen", lib.ptr(q.hi), lib.ptr(q.key), q.n, lib.ptr(y.hi),
                     lib.ptr(y.key), y.n, q.dpad, cap, steps, chained, lib.ptr(cand), lib.ptr(ckey),
                     lib.ptr(flag), lib.ptr(tau), lib.ptr(col_cnt), lib.ptr(col_buf), col_cap,
                     row_id_base, lib.stream_ptr())
        if prof is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            prof.append((ev0, ev1, q.n, y.n, q.d, "screen-dual" if dual is not None else "screen"))
        return cand, ckey, lists

    def _refine_checked(self, q: PreparedRows, y: PreparedRows, cand, k: int, exclude_self: bool,
                        tau, tau_row_stride: int, tau_step: int, tau_count: int, out=None):
        """Exact finish + completeness proof; returns (dist, ind, unverified int32 [nq]), written
        into the contiguous `out` triple when given."""
        lib = self._lib
        dev = self.device
        if out is not None:
            out_d, out_i, unverified = out
        else:
            out_d = torch.empty((q.n, k), dtype=torch.float64, device=dev)
            out_i = torch.empty((q.n, k), dtype=torch.int64, device=dev)
            unverified = torch.empty((q.n,), dtype=torch.int32, device=dev)
        lib.call("kb2_refine_topk_checked", lib.ptr(q.raw), q.n, q.raw.stride(0), lib.ptr(y.raw), y.n,
                 y.raw.stride(0), q.d, 4, lib.ptr(q.sqnorm), lib.ptr(y.sqnorm), lib.ptr(cand),
                 cand.shape[1], self._metric_code, y.base, int(exclude_self), y.base - q.base, k,
                 lib.ptr(out_d), lib.ptr(out_i), tau, tau_row_stride, tau_step, tau_count,
                 lib.ptr(q.key), lib.ptr(y.keymax), lib.ptr(q.err), lib.ptr(y.errmax),
                 self._eps_acc(q.dpad), lib.ptr(unverified), lib.stream_ptr())
        return out_d, out_i, unverified

    def _research(self, q: PreparedRows, y: PreparedRows, bad, k: int, exclude_self: bool,
                  out_d, out_i):
        """Rows `bad` (local query ids) whose proof failed: search them with the 3xTF32 kernel."""
        self.search_stats["screen_unverified"] += int(bad.numel())
        if bad.numel() == 0:
            return
        kk = k + 1 if exclude_self else k
        d_b, i_b = self._search_tf32x3(q.take(bad), y, kk, exclude_self=False)
        if exclude_self:
            # the gathered queries lost their row numbers: drop the own id here (it is there
            # unless duplicates of the row pushed it out, then drop the last)
            own = bad + q.base
            hit = i_b == own[:, None]
            drop = torch.where(hit.any(dim=1), hit.to(torch.int8).argmax(dim=1),
                               torch.full_like(own, kk - 1))
            pos = torch.arange(k, device=bad.device)[None, :]
            sel = pos + (pos >= drop[:, None]).to(pos.dtype)
            d_b, i_b = d_b.gather(1, sel), i_b.gather(1, sel)
        out_d[bad] = d_b
        out_i[bad] = i_b

    # -- dual-direction pass ------------------------------------------------------
    # The column thresholds start from a search of the columns against a strided SAMPLE of the
    # rows (rows / FUSED_SAMPLE_DIV of them) and tighten between row SEGMENTS whose sizes grow
    # geometrically (each segment ~ (FUSED_SEGMENT_GROWTH - 1) x the rows seen before it), so a
    # column receives ~cap (growth - 1) emits per segment: ~cap log2(DIV) in total at growth 2
    # instead of cap * DIV from one threshold.  Measured at C4 the pass costs ~1.2 ms per emitted
    # row per column (profiles/r01_ab_experiments.md block H).  KB2_FUSED_* override for tuning.
    FUSED_SAMPLE_DIV = float(os.environ.get("KB2_FUSED_SAMPLE_DIV", "32"))
    FUSED_SEGMENT_GROWTH = float(os.environ.get("KB2_FUSED_GROWTH", "3"))     # <= 1: one segment
    FUSED_SEGMENT_MIN_ROWS = int(os.environ.get("KB2_FUSED_MIN_ROWS", "16384"))
    FUSED_COL_CAP = int(os.environ.get("KB2_FUSED_COL_CAP", "512"))          # slots per column buffer
    SAMPLE_CHUNK_ROWS = 131072      # columns per threshold-search launch while the host upload runs
    FUSED_RANK_MIN_ROWS = 38144     # rows per rank and segment of a row-sharded pass (298 query tiles)

    def _use_fused(self, rows: PreparedRows, cols: PreparedRows, k: int) -> bool:
        if self.fused is False or self.impl in ("simt", "tc1"):
            return False
        cap = candidate_capacity(max(k, self.n_candidates))
        if cap > 64 or rows.d != cols.d or rows.n < 2 * cap or cols.n < 1:
            return False
        if self.fused is True:
            return True
        # auto: where the pass beats two one-direction passes (measured with the 1xTF32 screen:
        # C4 at c = 50 runs 1388 ms against 2165 ms, C5 -- d = 128, c = 50, 10 M columns --
        # 8.07 s against 10.13 s), and only for problems large enough to amortise the
        # threshold sample
        if rows.dpad < 128 or rows.n * cols.n < (1 << 32):
            return False
        # the column buffers must fit; decided from rank-independent quantities only (shapes,
        # world size, the device's TOTAL memory) so that every rank of a distributed run takes
        # the same branch -- the branches issue different collectives
        world = torch.distributed.get_world_size() if self.distributed else 1
        if self.distributed and self.shard_mode == "rows":
            world = 1                               # every rank buffers all columns
        total = torch.cuda.get_device_properties(self.device).total_memory
        return (cols.n // world + 1) * self._fused_col_cap(cap) * 8 < 0.3 * total

    def _fused_col_cap(self, cap: int) -> int:
        return max(self.FUSED_COL_CAP, cap)

    def _fused_sample_rows(self, n_rows: int, cap: int) -> int:
        return int(min(n_rows, max(8 * cap, -(-n_rows // max(1.0, self.FUSED_SAMPLE_DIV)))))

    def _fused_segments(self, n_rows: int, n_sample: int, min_rows: int = 0):
        """Row segment boundaries [0, b1, ..., n_rows] (multiples of 256 = one tile pair)."""
        g = self.FUSED_SEGMENT_GROWTH
        if g <= 1.0:
            return [0, n_rows]
        bounds, seen = [0], max(1, n_sample)
        while bounds[-1] < n_rows:
            size = max(int((g - 1.0) * seen), self.FUSED_SEGMENT_MIN_ROWS, min_rows, 256)
            nxt = (bounds[-1] + size + 255) // 256 * 256
            if n_rows - nxt < size // 2:        # fold a short tail into this segment
                nxt = n_rows
            bounds.append(min(n_rows, nxt))
            seen += bounds[-1] - bounds[-2]
        return bounds

    def search_both(self, rows: PreparedRows, cols: PreparedRows, k_rows: int, k_cols: int,
                    exclude_self_rows: bool = False, comm=None):
        """One contraction, both directions: ((dist, ind) of every row's k_rows nearest columns,
        (dist, ind) of every column's k_cols nearest rows).

        `comm` (distributed.RowShardComm): this rank holds a SHARD of the rows.  The row side
        is then complete per rank (every column is visited locally); the column side is agreed
        across ranks -- the thresholds through an all-gather of each column's best keys after
        the sample and between row segments, the result through an all-to-all of the column
        heads after the last one -- and each rank finishes a shard of the columns.  Returns
        (forward result of the own rows, reverse result of the own column shard)."""
        lib = self._lib
        dev = self.device
        cap = self._capacity(max(k_rows, k_cols), rows, dual=True)
        world = comm.world if comm is not None else 1
        n_total = comm.n_rows if comm is not None else rows.n
        with torch.cuda.device(dev):
            st = lib.stream_ptr()
            sm = torch.cuda.get_device_properties(dev).multi_processor_count
            screen = self._use_screen(rows, cols, cap, dual=True)
            prof = getattr(self, "_profile", None)       # bench.py: CUDA events around the searches
            # 1. column thresholds from a strided sample of the rows (a host matrix still being
            #    uploaded sent exactly these rows ahead, see _start_uploads); with sharded rows
            #    every rank searches its share of the sample
            n_s_total = self._fused_sample_rows(n_total, cap)
            n_s = n_s_total
            if comm is None:
                step = max(1, rows.n // n_s)
                if rows.presample is not None and rows.presample[0] == (n_s, step):
                    sample = rows.presample[1].ensure()
                else:
                    sample = rows.take(torch.arange(n_s, device=dev, dtype=torch.int64) * step)
                s_cols = cols
            else:
                # the sample comes from ALL rows (every rank holds them); each rank searches it
                # for a shard of the columns, so the pass costs 1 / world of the single-GPU one
                # and yields the very same thresholds
                step = max(1, n_total // n_s)
                sample = comm.rows_full.take(torch.arange(n_s, device=dev, dtype=torch.int64) * step)
                sc0, sc1, _per = comm.column_shard(cols.n)
                s_cols = cols.rows(sc0, sc1) if sc1 > sc0 else None
            if s_cols is None:
                tau = torch.empty((0,), dtype=torch.float32, device=dev)
                s_key = _s_idx = None
            elif screen and self._use_screen(s_cols, sample, cap, dual=False):
                if s_cols._pending is not None and s_cols.n >= 2 * self.SAMPLE_CHUNK_ROWS:
                    # the columns are still arriving from the host: search them chunk by chunk
                    # (each launch waits only for its own rows), the upload hides behind the search
                    tau = torch.empty((s_cols.n,), dtype=torch.float32, device=dev)
                    for c_lo in range(0, s_cols.n, self.SAMPLE_CHUNK_ROWS):
                        c_hi = min(s_cols.n, c_lo + self.SAMPLE_CHUNK_ROWS)
                        if s_cols.n - c_hi < self.SAMPLE_CHUNK_ROWS // 2:
                            c_hi = s_cols.n
                        _i, k_part, l_part = self._screen_search(s_cols.rows(c_lo, c_hi), sample, cap)
                        tau[c_lo:c_hi] = k_part.view(c_hi - c_lo, l_part, cap)[:, :, cap - 1].amin(dim=1)
                        del _i, k_part
                        if c_hi == s_cols.n:
                            break
                    s_key = _s_idx = None
                else:
                    s_cols.ensure()
                    _s_idx, s_key, lists = self._screen_search(s_cols, sample, cap)
            else:
                s_cols.ensure()
                lists = lib.lib.kb2_suggest_splits(s_cols.n, n_s, cap, sm)
                _s_idx = torch.empty((s_cols.n, lists * cap), dtype=torch.int32, device=dev)
                s_key = torch.empty((s_cols.n, lists * cap), dtype=torch.float32, device=dev)
                if prof is not None:
                    ev0 = torch.cuda.Event(enable_timing=True)
                    ev0.record()
                lib.call("kb2_knn_candidates", lib.KNN_AUTO, lib.ptr(s_cols.hi), lib.ptr(s_cols.lo),
                         s_cols.n, lib.ptr(sample.hi), lib.ptr(sample.lo), lib.ptr(sample.key), n_s,
                         rows.dpad, cap, lists, lib.ptr(_s_idx), lib.ptr(s_key), st)
                if prof is not None:
                    ev1 = torch.cuda.Event(enable_timing=True)
                    ev1.record()
                    prof.append((ev0, ev1, s_cols.n, n_s, rows.d, "tf32x3"))
            # the cap-th best within ANY subset of the rows bounds the final cap-th best
            if s_key is not None:
                tau = s_key.view(s_cols.n, lists, cap)[:, :, cap - 1].amin(dim=1).contiguous()
            if comm is not None:
                tau = comm.gather_columns(tau, cols.n)
            del _s_idx, s_key, sample
            cols.ensure()
            # 2. the dual-direction pass, one launch per row segment; exact finish of the row
            #    lists per segment; thresholds tighten between segments
            col_cap = self._fused_col_cap(cap)
            col_cnt = torch.zeros(cols.n, dtype=torch.int32, device=dev)
            col_buf = torch.empty((cols.n, col_cap), dtype=torch.int64, device=dev)
            fwd_d = torch.empty((rows.n, k_rows), dtype=torch.float64, device=dev)
            fwd_i = torch.empty((rows.n, k_rows), dtype=torch.int64, device=dev)
            unv_rows = torch.zeros((rows.n,), dtype=torch.int32, device=dev) if screen else None
            # sharded rows: every rank's share of a segment should still fill the GPU with chained
            # query tiles (>= 2 tiles per SM), else the launch falls back to independent index
            # ranges, whose lists each pay the fill phase again
            bounds = self._fused_segments(
                n_total, n_s_total,
                min_rows=0 if comm is None else min(world * self.FUSED_RANK_MIN_ROWS, n_total // 2))
            if comm is not None:
                # the same number of segments on every rank (the exchanges between them are
                # collective), each a proportional share of the global segment
                bounds = [min(rows.n, (int(b * rows.n / n_total) + 255) // 256 * 256) for b in bounds]
                bounds[0], bounds[-1] = 0, rows.n
                bounds = [max(b, p) for b, p in zip(bounds, [0] + bounds[:-1])]
            # statistics for bench.py / tests (extra reductions and host syncs): only on request
            stats = bool(getattr(self, "_collect_stats", False))
            emitted = torch.zeros((), dtype=torch.int64, device=dev) if stats else None
            probe = None
            for s_no, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
                last = s_no == len(bounds) - 2
                seg = rows.rows(lo, hi) if hi > lo else None
                if seg is None:
                    pass                                    # a rank without rows in this segment
                elif screen:
                    cand_rows, key_rows, r_lists = self._screen_search(
                        seg, cols, cap, dual=(tau, col_cnt, col_buf, col_cap, lo))
                    self._refine_checked(seg, cols, cand_rows, k_rows, exclude_self_rows,
                                         lib.ptr(key_rows) + 4 * (cap - 1), r_lists * cap, cap, r_lists,
                                         out=(fwd_d[lo:hi], fwd_i[lo:hi], unv_rows[lo:hi]))
                    if s_no == 0 and self._screen_ok is None:
                        probe = self._probe_fraction(unv_rows[lo:hi], seg, cols, key_rows, r_lists, cap,
                                                     fwd_d[lo:hi], k_rows)
                    del key_rows, cand_rows
                else:
                    splits = lib.lib.kb2_suggest_splits(seg.n, cols.n, cap, sm)
                    cand_rows = torch.empty((seg.n, splits * cap), dtype=torch.int32, device=dev)
                    if prof is not None:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record()
                    lib.call("kb2_knn_fused", lib.ptr(seg.hi), lib.ptr(seg.lo), lib.ptr(seg.key), seg.n,
                             lib.ptr(cols.hi), lib.ptr(cols.lo), lib.ptr(cols.key), cols.n, rows.dpad,
                             cap, splits, lib.ptr(tau), lib.ptr(col_cnt), lib.ptr(col_buf), col_cap,
                             lo, lib.ptr(cand_rows), st)
                    if prof is not None:
                        ev1 = torch.cuda.Event(enable_timing=True)
                        ev1.record()
                        prof.append((ev0, ev1, seg.n, cols.n, rows.d, "tf32x3-dual"))
                    self._refine(seg, cols, cand_rows, k_rows, exclude_self_rows,
                                 out=(fwd_d[lo:hi], fwd_i[lo:hi]))
                    del cand_rows
                if screen and s_no == 0 and len(bounds) > 2 and not self._screen_verdict(
                        probe if seg is not None else None, cap, rows.dpad, dual=True, comm=comm):
                    # probe failed: start over with longer lists, or with 3xTF32 keys
                    # (_capacity / _use_screen now answer differently)
                    del col_buf, col_cnt, fwd_d, fwd_i, unv_rows, tau
                    return self.search_both(rows, cols, k_rows, k_cols,
                                            exclude_self_rows=exclude_self_rows, comm=comm)
                if not last or comm is not None:
                    # best cap rows to the head of every column buffer, tau = cap-th best so far
                    if emitted is not None:     # sticky overflow counts are not emits
                        emitted += torch.where(col_cnt < (1 << 30), col_cnt, 0).sum()
                    lib.call("kb2_col_compact", lib.ptr(col_buf), lib.ptr(col_cnt), cols.n, col_cap,
                             cap, lib.ptr(tau), st)
                    if emitted is not None:
                        emitted -= torch.where(col_cnt < (1 << 30), col_cnt, 0).sum()
                    if comm is not None and not last:
                        # ... over the rows of ALL ranks: exchange the heads' keys
                        keys = torch.empty((cols.n, cap), dtype=torch.float32, device=dev)
                        lib.call("kb2_col_heads", lib.ptr(col_buf), lib.ptr(col_cnt), cols.n, col_cap,
                                 cap, 0, None, lib.ptr(keys), st)
                        comm.kth_over_ranks(keys, cap, tau=tau)
                        del keys
            rows.ensure()
            bad_rows = torch.zeros(0, dtype=torch.int64, device=dev)
            if screen:
                self.search_stats["screen_rows"] += rows.n
                bad_rows = torch.nonzero(unv_rows).flatten()
                self._research(rows, cols, bad_rows, k_rows, exclude_self_rows, fwd_d, fwd_i)
            fwd = (fwd_d, fwd_i)
            # 3. column side: best cap emitted rows per column, exact finish
            if comm is None:
                q_cols, y_rows, c0 = cols, rows, 0
                cand_cols = torch.empty((cols.n, cap), dtype=torch.int32, device=dev)
                overflow = torch.empty(cols.n, dtype=torch.int32, device=dev)
                col_tau = torch.empty(cols.n, dtype=torch.float32, device=dev) if screen else None
                lib.call("kb2_col_select", lib.ptr(col_buf), lib.ptr(col_cnt), cols.n, col_cap, cap,
                         lib.ptr(cand_cols), lib.ptr(overflow), lib.ptr(tau) if screen else None,
                         lib.ptr(col_tau), st)
            else:
                # every rank holds the best cap of ITS rows per column (heads after the final
                # compaction) and a bound tau for everything it did not keep; the owner of a
                # column shard merges the W heads and finishes against ALL rows
                heads = torch.empty((cols.n, cap), dtype=torch.int64, device=dev)
                lib.call("kb2_col_heads", lib.ptr(col_buf), lib.ptr(col_cnt), cols.n, col_cap, cap,
                         rows.base, lib.ptr(heads), None, st)
                lost = ((col_cnt >= (1 << 30)) | (col_cnt > col_cap)).to(torch.int32)
                comm.max_(lost)
                comm.min_(tau)
                recv, c0, c1 = comm.columns_to_owners(heads, lib.EMPTY_ENTRY)
                del heads
                m_loc = c1 - c0
                q_cols, y_rows = cols.rows(c0, c1), comm.rows_full
                buf2 = recv.permute(1, 0, 2)[:m_loc].reshape(m_loc, world * cap).contiguous()
                cnt2 = torch.full((m_loc,), world * cap, dtype=torch.int32, device=dev)
                cand_cols = torch.empty((m_loc, cap), dtype=torch.int32, device=dev)
                overflow = torch.empty(m_loc, dtype=torch.int32, device=dev)
                tau_loc = tau[c0:c1].contiguous()
                col_tau = torch.empty(m_loc, dtype=torch.float32, device=dev)
                if m_loc:
                    lib.call("kb2_col_select", lib.ptr(buf2), lib.ptr(cnt2), m_loc, world * cap, cap,
                             lib.ptr(cand_cols), lib.ptr(overflow), lib.ptr(tau_loc), lib.ptr(col_tau), st)
                # what the merge dropped is bounded by its cap-th key, what a rank never kept by
                # that rank's bound
                col_tau = torch.minimum(col_tau, tau_loc)
                overflow = lost[c0:c1].contiguous()
                del recv, buf2
            del col_buf
            if screen:
                rev_d, rev_i, unv = self._refine_checked(q_cols, y_rows, cand_cols, k_cols, False,
                                                         lib.ptr(col_tau), 1, 1, 1)
                self.search_stats["screen_rows"] += q_cols.n
                n_over = int(overflow.sum()) if stats else 0
                # overflowed columns lost rows below their threshold: search them again too
                bad_cols = torch.nonzero(unv | overflow).flatten()
                self._research(q_cols, y_rows, bad_cols, k_cols, False, rev_d, rev_i)
                # global ids of the rows / columns that took the 3xTF32 re-search (parity samples
                # of the tests and bench.py force-include them)
                self.researched = {"rows": bad_rows + rows.base, "cols": bad_cols + q_cols.base}
            else:
                rev_d, rev_i = self._refine(q_cols, y_rows, cand_cols, k_cols, False)
                # overflowed columns lost rows; a column whose threshold TIES with its cap-th
                # best key (duplicate rows: emits are strictly below tau) may hold fewer than k
                bad = torch.nonzero(overflow.bool() | (rev_i[:, k_cols - 1] < 0)).flatten()
                n_over = int(bad.numel())
                if bad.numel():      # columns whose buffer overflowed: plain search for those few
                    d_b, i_b = self._search_tf32x3(q_cols.take(bad), y_rows, k_cols)
                    rev_d[bad] = d_b
                    rev_i[bad] = i_b
            if stats:
                emitted += torch.where(col_cnt < (1 << 30), col_cnt, 0).sum()
                self._fused_stats = {"sample_rows": int(n_s), "col_cap": int(col_cap),
                                     "row_segments": [int(b) for b in bounds],
                                     "emitted_per_column_mean": float(emitted) / max(1, cols.n),
                                     "emitted_scope": "this rank's rows" if comm is not None else "all rows",
                                     "overflow_columns": n_over}
        return fwd, (rev_d, rev_i)

    def _refine(self, q: PreparedRows, y: PreparedRows, cand, k: int, exclude_self: bool, out=None):
        """Exact float64 finish of candidate lists `cand` (local ids into y)."""
        lib = self._lib
        dev = self.device
        if out is not None:
            out_d, out_i = out
        else:
            out_d = torch.empty((q.n, k), dtype=torch.float64, device=dev)
            out_i = torch.empty((q.n, k), dtype=torch.int64, device=dev)
        q_raw, y_raw = q.raw, y.raw
        if q_raw.dtype != y_raw.dtype:
            q_raw, y_raw = q_raw.to(torch.float64), y_raw.to(torch.float64)
        # fewer than k valid candidates (a shard smaller than k) come back as +inf / -1
        lib.call("kb2_refine_topk", lib.ptr(q_raw), q.n, q_raw.stride(0), lib.ptr(y_raw), y.n,
                 y_raw.stride(0), q.d, q_raw.element_size(), lib.ptr(q.sqnorm), lib.ptr(y.sqnorm),
                 lib.ptr(cand), cand.shape[1], self._metric_code, y.base, int(exclude_self),
                 y.base - q.base, k, lib.ptr(out_d), lib.ptr(out_i), lib.stream_ptr())
        return out_d, out_i

    def search(self, q: PreparedRows, y: PreparedRows, k: int, exclude_self: bool = False,
               splits: Optional[int] = None):
        """k nearest rows of `y` for every row of `q`: (dist float64, ind int64) on device,
        rows ascending, ids global (y.base added)."""
        lib = self._lib
        if q.d != y.d:
            raise ValueError(f"query has {q.d} features, index has {y.d}")
        q.ensure()
        y.ensure()
        cap = min(self._capacity(k, q), lib.lib.kb2_max_candidates())
        if k > cap:
            raise ValueError(
                f"B200 supports at most {lib.lib.kb2_max_candidates()} neighbours per "
                f"query and shard, got {k}")
        if q.n == 0 or splits is not None or not self._use_screen(q, y, cap, dual=False) \
                or (exclude_self and y.n < k + 2):
            return self._search_tf32x3(q, y, k, exclude_self=exclude_self, splits=splits)
        dev = self.device
        with torch.cuda.device(dev):
            out_d = torch.empty((q.n, k), dtype=torch.float64, device=dev)
            out_i = torch.empty((q.n, k), dtype=torch.int64, device=dev)
            unv = torch.empty((q.n,), dtype=torch.int32, device=dev)
            parts = [(0, q.n)]
            if self.precision == "auto" and self._screen_ok is None and \
                    q.n >= 4 * self.SCREEN_PROBE_ROWS:
                parts = [(0, self.SCREEN_PROBE_ROWS), (self.SCREEN_PROBE_ROWS, q.n)]
            for lo, hi in parts:
                part = q if (lo, hi) == (0, q.n) else q.rows(lo, hi)
                if self._screen_ok is False:          # the probe said no: 3xTF32 for the rest
                    d_p, i_p = self._search_tf32x3(part, y, k, exclude_self=exclude_self)
                    out_d[lo:hi] = d_p
                    out_i[lo:hi] = i_p
                    unv[lo:hi] = 0
                    continue
                cand, ckey, lists = self._screen_search(part, y, cap)
                self._refine_checked(part, y, cand, k, exclude_self, lib.ptr(ckey) + 4 * (cap - 1),
                                     lists * cap, cap, lists,
                                     out=(out_d[lo:hi], out_i[lo:hi], unv[lo:hi]))
                probe = None
                if len(parts) > 1 and lo == 0 and self._screen_ok is None:
                    probe = self._probe_fraction(unv[lo:hi], part, y, ckey, lists, cap, out_d[lo:hi], k)
                del cand, ckey
                if len(parts) > 1 and lo == 0 and not self._screen_verdict(probe, cap, q.dpad) \
                        and self._screen_ok is None:
                    # the probe asks for longer lists: start over (3xTF32 verdicts carry on below)
                    del out_d, out_i, unv
                    return self.search(q, y, k, exclude_self=exclude_self)
                self.search_stats["screen_rows"] += hi - lo
            self._research(q, y, torch.nonzero(unv).flatten(), k, exclude_self, out_d, out_i)
        return out_d, out_i

    def _search_tf32x3(self, q: PreparedRows, y: PreparedRows, k: int, exclude_self: bool = False,
                       splits: Optional[int] = None):
        """The 3xTF32 search (knn_tc2.cu / knn_tc.cu / the SIMT cross-check) + exact finish."""
        lib = self._lib
        dev = self.device
        q.ensure()
        y.ensure()
        with torch.cuda.device(dev):
            out_d = torch.empty((q.n, k), dtype=torch.float64, device=dev)
            out_i = torch.empty((q.n, k), dtype=torch.int64, device=dev)
            if q.n == 0:
                return out_d, out_i
            cap = min(candidate_capacity(k), lib.lib.kb2_max_candidates())
            if splits is None:
                sm = torch.cuda.get_device_properties(dev).multi_processor_count
                splits = lib.lib.kb2_suggest_splits(q.n, y.n, cap, sm)
            ncand = splits * cap
            cand = torch.empty((q.n, ncand), dtype=torch.int32, device=dev)
            impl = {"auto": lib.KNN_AUTO, "tc": lib.KNN_TC, "simt": lib.KNN_SIMT,
                    "tc1": lib.KNN_TC1}[self.impl]
            st = lib.stream_ptr()
            # self = index row j with j + self_offset == query row (ids local to q / y); the
            # search keeps it (one margin slot), the exact finish drops it by id
            prof = getattr(self, "_profile", None)   # bench.py: CUDA events around the search
            if prof is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            lib.call("kb2_knn_candidates", impl, lib.ptr(q.hi), lib.ptr(q.lo), q.n,
                     lib.ptr(y.hi), lib.ptr(y.lo), lib.ptr(y.key), y.n, q.dpad, cap, splits,
                     lib.ptr(cand), None, st)
            if prof is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                prof.append((ev0, ev1, q.n, y.n, q.d, "tf32x3"))
            out_d, out_i = self._refine(q, y, cand, k, exclude_self)
        return out_d, out_i


class B200(B200Mixin, NNAlgorithm):
    """Exact kNN on one or more B200 GPUs; ``Kiez(algorithm="B200")``.

    Parameters mirror SklearnNN where they apply (n_candidates, metric, p, n_jobs).
    ``impl``: "auto"/"tc" = tcgen05 tensor-core search (CTA pairs), "tc1" = its single-CTA
    form, "simt" = FP32-pipe cross-check.
    ``precision``: "tf32x3" = every candidate key from the 3xTF32 split; "screen" / "auto" =
    propose candidates with ONE TF32 product (3x fewer MMAs, knn_screen.cu), prove in the
    float64 finish that the proposal contains the exact top k, search the few rows where the
    proof fails again with 3xTF32 -- same results (``search_stats`` counts the fallbacks).
    With "auto" the first rows of a fit are a probe: if the proof fails for more than
    ``SCREEN_MAX_UNVERIFIED`` of them (data whose neighbour gaps are below the TF32 error
    bound) the rest of the fit uses the 3xTF32 kernels; "screen" never falls back.
    ``fused``: "auto" / True / False -- serve kiez's reverse and forward kNN from ONE
    dual-direction pass (automatic for cap <= 32, d >= 192, n*m >= 2^32).
    ``distributed``: shard the index side over the ranks of an initialised
    torch.distributed (NCCL) group; default: on when world_size > 1.
    """

    if torch is not None:
        _ALLOWED_INPUT_TYPES = (np.ndarray, torch.Tensor)
