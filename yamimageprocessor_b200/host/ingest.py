"""Host <-> device streaming for frames that do not fit a single staging copy.

The reference pages huge mosaics through ``np.memmap`` row slices (``core/tiled_image.py:85-96,
134-157``, ``core/io_manager.py:234-242``).  ``upload_rows`` moves a band of rows of such a source
into device memory through a small ring of page-locked buffers: while the DMA of chunk k runs on a
dedicated copy stream, host threads gather chunk k+1 (page-cache reads / page faults of the memmap
happen there, in parallel).  ``download_into`` is the mirror image for results that go back into a
caller-owned (pageable) array.  Both keep the ring per backend, so nothing is re-pinned per call.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np

from ..backend import Backend

_CHUNK_BYTES = 64 << 20
_DEPTH = 3
_POOL: Optional[ThreadPoolExecutor] = None
_WORKERS = max(2, min(16, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))))


def _pool() -> ThreadPoolExecutor:
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=_WORKERS, thread_name_prefix="yam-ingest")
    return _POOL


def _parallel_copy(dst: np.ndarray, src) -> None:
    """dst[...] = src with the rows split over the pool (np.copyto releases the GIL)."""
    rows = dst.shape[0]
    parts = min(rows, _WORKERS)
    if parts <= 1 or dst.nbytes < (4 << 20):
        np.copyto(dst, src)
        return
    bounds = [rows * i // parts for i in range(parts + 1)]
    list(_pool().map(lambda i: np.copyto(dst[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]]), range(parts)))


class _Ring:
    def __init__(self, chunk_bytes: int, depth: int):
        import torch

        self.chunk_bytes = chunk_bytes
        self.bufs = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
        self.events = [None] * depth


def _ring(be: Backend, row_bytes: int) -> _Ring:
    chunk = max(_CHUNK_BYTES, row_bytes)
    ring = getattr(be, "_ingest_ring", None)
    if ring is None or ring.chunk_bytes < chunk:
        ring = _Ring(chunk, _DEPTH)
        be._ingest_ring = ring  # type: ignore[attr-defined]
    return ring


def _copy_stream(be: Backend):
    import torch

    s = getattr(be, "_ingest_stream", None)
    if s is None:
        s = torch.cuda.Stream(device=be.device)
        be._ingest_stream = s  # type: ignore[attr-defined]
    return s


def upload_rows(be: Backend, source, r0: int, r1: int):
    """Rows ``[r0, r1)`` of a 2-D host array / memmap -> new CUDA tensor ``(r1 - r0, W)``.

    Asynchronous with respect to the caller's stream: the returned tensor may be used right away on
    the current stream (it waits for the copy stream).
    """
    import torch

    if getattr(source, "ndim", 0) != 2:
        raise ValueError("upload_rows expects a 2-D (H, W) source")
    H, W = int(source.shape[0]), int(source.shape[1])
    if not (0 <= r0 < r1 <= H):
        raise ValueError(f"row range [{r0}, {r1}) outside the source ({H} rows)")
    dt = np.dtype(source.dtype)
    if dt not in be._np2t() or dt == np.int64:
        raise TypeError(f"unsupported dtype {dt}; expected uint8, uint16, float32 or int32")
    rows = r1 - r0
    row_bytes = W * dt.itemsize
    out = torch.empty((rows, W), dtype=be._np2t()[dt], device=be.device)
    flat = out.view(torch.uint8).reshape(-1)
    stream = _copy_stream(be)
    with be._lock:  # one user of the ring at a time
        ring = _ring(be, row_bytes)
        rows_per_chunk = max(1, ring.chunk_bytes // row_bytes)
        stream.wait_stream(torch.cuda.current_stream(be.device))  # `out` was allocated on the current stream
        for k, b0 in enumerate(range(0, rows, rows_per_chunk)):
            b1 = min(rows, b0 + rows_per_chunk)
            slot = k % len(ring.bufs)
            if ring.events[slot] is not None:
                ring.events[slot].synchronize()  # the DMA that last read this slot has finished
            nbytes = (b1 - b0) * row_bytes
            host = ring.bufs[slot].numpy()[:nbytes].view(dt).reshape(b1 - b0, W)
            _parallel_copy(host, source[r0 + b0:r0 + b1])
            with torch.cuda.stream(stream):
                flat[b0 * row_bytes:b0 * row_bytes + nbytes].copy_(ring.bufs[slot][:nbytes], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            ring.events[slot] = ev
        torch.cuda.current_stream(be.device).wait_stream(stream)
    return out


def download_into(be: Backend, t, out: np.ndarray) -> np.ndarray:
    """CUDA tensor -> ``out`` (a C-contiguous host array of the same shape and dtype; pageable is
    fine).  The host-side copy of chunk k overlaps the DMA of chunk k+1.  Synchronous."""
    import torch

    t = t.contiguous()
    if tuple(out.shape) != tuple(t.shape) or be._np2t().get(np.dtype(out.dtype)) != t.dtype or not out.flags.c_contiguous:
        raise ValueError("download_into: `out` must be C-contiguous with the tensor's shape and dtype")
    total = out.nbytes
    if total == 0:
        return out
    flat = t.view(torch.uint8).reshape(-1)
    dst = out.reshape(-1).view(np.uint8)
    stream = _copy_stream(be)
    with be._lock:
        ring = _ring(be, 1)
        chunk = ring.chunk_bytes
        stream.wait_stream(torch.cuda.current_stream(be.device))
        pending = []  # (event, slot, offset, nbytes) of DMAs in flight

        def drain(item):
            ev, slot, off, nbytes = item
            ev.synchronize()
            src = ring.bufs[slot].numpy()[:nbytes]
            piece = dst[off:off + nbytes]
            parts = min(_WORKERS, max(1, nbytes >> 22))
            if parts <= 1:
                np.copyto(piece, src)
            else:
                bounds = [nbytes * i // parts for i in range(parts + 1)]
                list(_pool().map(lambda i: np.copyto(piece[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]]), range(parts)))
            ring.events[slot] = None

        for k, off in enumerate(range(0, total, chunk)):
            slot = k % len(ring.bufs)
            if len(pending) == len(ring.bufs):
                drain(pending.pop(0))  # frees the slot this iteration reuses
            elif ring.events[slot] is not None:
                ring.events[slot].synchronize()
            nbytes = min(chunk, total - off)
            with torch.cuda.stream(stream):
                ring.bufs[slot][:nbytes].copy_(flat[off:off + nbytes], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            ring.events[slot] = ev
            pending.append((ev, slot, off, nbytes))
        while pending:
            drain(pending.pop(0))
        torch.cuda.current_stream(be.device).wait_stream(stream)
    return out


__all__ = ["download_into", "upload_rows"]
