"""Host-to-device upload of an embedding matrix that OVERLAPS with the search.

kiez hands the backend host arrays (numpy, pageable): `NNAlgorithm.fit` keeps references
(neighbor_algorithm_base.py:93-96) and the searches read them.  A plain `tensor.to(device)` of
two 1 GB matrices costs more than a tenth of a C4 step and nothing else runs meanwhile.  Here
one background thread per `fit` moves each matrix in row chunks: pageable rows are copied
into a small ring of pinned staging buffers (the memcpy runs on torch's intra-op threads and
releases the GIL), every chunk goes to the device with an asynchronous copy on a dedicated copy
stream and is published with a CUDA event.  The consumer (`PreparedRows.ensure`) waits for the
chunks it needs -- on the host until the copy has been enqueued, on its stream for the copy
itself -- so the dual-direction pass starts on the first rows while the rest is still in flight.
Pinned input (numpy views of pinned torch tensors) skips the staging ring.

torch is used for what it is here for: device memory, streams, events.
"""
from __future__ import annotations

import os
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Tuple

import numpy as np
import torch

CHUNK_BYTES = 32 << 20          # one staging buffer / one async copy
RING = 3                        # staging buffers in flight
# the pageable -> pinned memcpy of a chunk is split over this many helper threads (numpy / torch
# copies release the GIL): torchrun exports OMP_NUM_THREADS=1 to every rank, which would leave a
# single-threaded memcpy (~6 GB/s) as the bottleneck of the upload.  Half of the cores this
# process may use, at most 8 (KB2_COPY_THREADS overrides)
def _usable_cpus() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):  # pragma: no cover
        return os.cpu_count() or 4


COPY_THREADS = int(os.environ.get("KB2_COPY_THREADS", "0")) or max(
    1, min(8, _usable_cpus() // 2 // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))

_staging = {}                   # (device index, nbytes) -> ([pinned uint8 tensors], [last copy event])
_copy_streams = {}              # device index -> torch.cuda.Stream
_upload_lock = threading.Lock()  # one upload thread at a time shares the staging ring


def _copy_stream(device) -> "torch.cuda.Stream":
    idx = torch.device(device).index or 0
    if idx not in _copy_streams:
        _copy_streams[idx] = torch.cuda.Stream(device=device)
    return _copy_streams[idx]


def _staging_ring(device, nbytes: int):
    """(buffers, events): events[i] = the last asynchronous copy that read buffers[i]."""
    key = (torch.device(device).index or 0, nbytes)
    if key not in _staging:
        _staging[key] = ([torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(RING)],
                         [None] * RING)
    return _staging[key]


_pool = None


def _copy_pool():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=COPY_THREADS, thread_name_prefix="kiez_b200-memcpy")
    return _pool


def eligible(data) -> bool:
    """Host fp32 row-major matrices take the overlapped path; everything else (float64 callers,
    strided views, device tensors) keeps the plain synchronous upload."""
    if isinstance(data, np.ndarray):
        return data.ndim == 2 and data.dtype == np.float32 and data.flags.c_contiguous and data.size > 0
    return (torch.is_tensor(data) and not data.is_cuda and data.dim() == 2
            and data.dtype == torch.float32 and data.is_contiguous() and data.numel() > 0)


class HostUpload:
    """One matrix (or a strided row sample of it) on its way to the device."""

    def __init__(self, data, device, rows: Optional[Tuple[int, int]] = None):
        host = torch.from_numpy(data) if isinstance(data, np.ndarray) else data
        if rows is not None:                       # (count, step): rows 0, step, 2 step, ...
            count, step = rows
            self._gather = (count, step)
            n = count
        else:
            self._gather = None
            n = host.shape[0]
        self.host = host
        self.n, self.d = n, host.shape[1]
        self.device = torch.device(device)
        self.dev = torch.empty((n, self.d), dtype=torch.float32, device=self.device)
        self.chunk_rows = max(256, (CHUNK_BYTES // (self.d * 4)) // 256 * 256)
        self.bounds = [(lo, min(n, lo + self.chunk_rows)) for lo in range(0, n, self.chunk_rows)]
        self.events: List[Optional[torch.cuda.Event]] = [None] * len(self.bounds)
        self.ready = [threading.Event() for _ in self.bounds]
        self.error: Optional[BaseException] = None
        self.t_enqueued = [0.0] * len(self.bounds)   # time.perf_counter() when chunk i was enqueued
        # the device buffer may be a recycled block that kernels of the consumer's stream still
        # read: the first copy waits for everything enqueued there so far
        self._alloc_event = torch.cuda.Event()
        self._alloc_event.record(torch.cuda.current_stream(self.device))

    def chunk_of(self, row: int) -> int:
        return min(len(self.bounds) - 1, row // self.chunk_rows)

    def wait(self, chunk: int, stream) -> None:
        """Block the host until chunk `chunk` has been enqueued, then make `stream` wait for it."""
        self.ready[chunk].wait()
        if self.error is not None:
            raise RuntimeError("host-to-device upload failed") from self.error
        stream.wait_event(self.events[chunk])

    # -- uploader thread -------------------------------------------------------------
    def run(self) -> None:
        try:
            torch.cuda.set_device(self.device)
            stream = _copy_stream(self.device)
            stream.wait_event(self._alloc_event)
            pinned = self._gather is None and self.host.is_pinned()
            ring, ring_events = (None, None) if pinned else \
                _staging_ring(self.device, self.chunk_rows * self.d * 4)
            pool = _copy_pool() if not pinned and COPY_THREADS > 1 else None
            for i, (lo, hi) in enumerate(self.bounds):
                if pinned:
                    src = self.host[lo:hi]
                else:
                    slot = i % RING
                    if ring_events[slot] is not None:
                        ring_events[slot].synchronize()      # the copy that used this buffer is done
                    buf = ring[slot][: (hi - lo) * self.d * 4].view(torch.float32).view(hi - lo, self.d)
                    if self._gather is None:
                        part = self.host[lo:hi]
                    else:
                        step = self._gather[1]
                        part = self.host[lo * step: (hi - 1) * step + 1: step]
                    if pool is None or hi - lo < 4 * COPY_THREADS:
                        buf.copy_(part)
                    else:
                        cuts = [(hi - lo) * t // COPY_THREADS for t in range(COPY_THREADS + 1)]
                        list(pool.map(lambda ab: buf[ab[0]:ab[1]].copy_(part[ab[0]:ab[1]]),
                                      zip(cuts[:-1], cuts[1:])))
                    src = buf
                with torch.cuda.stream(stream):
                    self.dev[lo:hi].copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                if not pinned:
                    ring_events[i % RING] = ev
                self.events[i] = ev
                self.t_enqueued[i] = time.perf_counter()
                self.ready[i].set()
        except BaseException as exc:  # surfaced by wait()
            self.error = exc
            for r in self.ready:
                r.set()


class Uploader:
    """Runs the uploads of one `fit` in order on one background thread."""

    def __init__(self):
        self.jobs: List[HostUpload] = []
        self._thread: Optional[threading.Thread] = None

    def add(self, job: HostUpload) -> HostUpload:
        self.jobs.append(job)
        return job

    def start(self) -> None:
        jobs = list(self.jobs)

        def work():
            with _upload_lock:
                for job in jobs:
                    job.run()

        self._thread = threading.Thread(target=work, name="kiez_b200-upload", daemon=True)
        self._thread.start()

    def join(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None
