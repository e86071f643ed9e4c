"""Device-resident tier of the reference's ``PipelineCache``.

``processing/pipeline_cache.py`` memoises every step output under a signature chain
(``predict:291-313``) and, per step, copies the array two or three times, ``np.save``s it and
fsyncs (``_compute_dense:352-394``).  This class keeps the same keys and the same
``register_source / predict / compute / discard_cache`` surface but stores the per-step outputs as
CUDA tensors: a re-run after a parameter change restarts from the last step whose signature still
matches, entirely on the device, and only the final image is downloaded.  Entries are evicted
least-recently-used beyond ``max_bytes``.  Steps run through ``B200Executor`` by NAME (the names
that feed the signatures); unknown names raise ``KeyError`` — there is no CPU fallback.

Progressive results (``PipelineCacheTileUpdate``, reference ``:91-105`` / ``_compute_tiled:416-574``):
with ``incremental=callback`` the final image is handed to the caller tile by tile, in the
reference's row-major box order (``core/tiled_image.py:15-30``), while the download of the following
band of tile rows is still in flight -- the UI's progressive preview path.  Unlike the reference's
tiled compute the tiles are cut from the DENSE result, so neighbourhood filters see their halos and
global statistics stay global.  A lazy ``TiledPipelineImage`` source is
uploaded by row bands straight from its handle (never densified on the host).

Disk tier: with ``cache_directory`` the FINAL image of every compute is written through in the
reference's layout and naming -- ``{source_id}_{signature}.npy``, or ``.npz`` with ``tile_{i}`` arrays
and a JSON ``metadata`` record of type ``"tiles"`` when tiles were requested (``:721-799``), atomically
(tmp -> fsync -> os.replace) -- and ``get_cached_image`` reloads it after eviction or in a new process.
"""
from __future__ import annotations

import contextlib
import hashlib
import json
import os
import threading
from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import cache_keys
from .cache_keys import StepRecord
from .executor import B200Executor
from .tiles import TiledPipelineImage, iter_tile_boxes

TileBox = Tuple[int, int, int, int]


@dataclass(frozen=True)
class PipelineCacheTileUpdate:
    """Incremental update emitted while results stream back (reference: processing/pipeline_cache.py:91-105;
    same fields, same meaning)."""

    source_id: str
    final_signature: str
    step_signature: str
    step_index: int
    total_steps: int
    box: TileBox
    tile: np.ndarray
    shape: Tuple[int, ...]
    dtype: np.dtype
    tile_size: Optional[Tuple[int, int]]
    from_cache: bool = False


class OperationCancelled(RuntimeError):
    """Raised when ``cancel_event`` is set between two steps (reference: checked per step, ``:371``)."""


@dataclass
class DeviceCacheResult:
    source_id: str
    final_signature: str
    image: np.ndarray            # fresh host copy of the final image
    steps: List[StepRecord]
    metadata: Dict[str, Any]
    device_image: Any = None     # the same image as a CUDA tensor (stays cached)
    computed: Tuple[int, ...] = ()  # indices of the steps that actually ran (the rest were cache hits)


class DevicePipelineCache:
    def __init__(self, executor: Optional[B200Executor] = None, max_bytes: int = 16 << 30,
                 cache_directory: Optional[os.PathLike] = None) -> None:
        self._executor = executor or B200Executor()
        self._max_bytes = int(max_bytes)
        self._cache_directory = Path(cache_directory) if cache_directory is not None else None
        if self._cache_directory is not None:
            self._cache_directory.mkdir(parents=True, exist_ok=True)
        self._lock = threading.RLock()
        self._entries: "OrderedDict[Tuple[str, str], Any]" = OrderedDict()  # (source_id, signature) -> tensor
        self._bytes = 0
        self._metadata: Dict[str, Dict[str, Dict[str, Any]]] = {}

    # ------------------------------------------------------------------ bookkeeping
    @staticmethod
    def _nbytes(t) -> int:
        return int(t.numel()) * int(t.element_size())

    def _put(self, source_id: str, signature: str, tensor) -> None:
        key = (source_id, signature)
        with self._lock:
            old = self._entries.pop(key, None)
            if old is not None:
                self._bytes -= self._nbytes(old)
            self._entries[key] = tensor
            self._bytes += self._nbytes(tensor)
            while self._bytes > self._max_bytes and len(self._entries) > 1:
                k, victim = next(iter(self._entries.items()))
                if k == key:
                    break
                self._entries.pop(k)
                self._bytes -= self._nbytes(victim)

    def _get(self, source_id: str, signature: str):
        key = (source_id, signature)
        with self._lock:
            t = self._entries.get(key)
            if t is not None:
                self._entries.move_to_end(key)
            return t

    @property
    def resident_bytes(self) -> int:
        return self._bytes

    # ------------------------------------------------------------------ reference surface
    def register_source(self, image: np.ndarray, *, hint: Optional[str] = None) -> str:
        """sha256(shape, dtype, bytes) like ``register_source:256-264``; the image is uploaded once."""
        sid = cache_keys.source_id(image)
        if self._get(sid, sid) is None:
            self._put(sid, sid, self._executor.backend.to_device(np.ascontiguousarray(image)))
        meta = {"version": 1, "source_id": sid, "final_signature": sid, "steps": []}
        if hint:
            meta["hint"] = str(hint)
        with self._lock:
            self._metadata.setdefault(sid, {})[sid] = meta
        return sid

    def discard_cache(self, source_id: str) -> None:
        with self._lock:
            for key in [k for k in self._entries if k[0] == source_id]:
                self._bytes -= self._nbytes(self._entries.pop(key))

    def predict(self, source_id: str, steps: Sequence[Any]) -> Tuple[str, List[StepRecord]]:
        return cache_keys.predict(source_id, steps)

    def metadata(self, source_id: str, final_signature: str) -> Optional[Dict[str, Any]]:
        with self._lock:
            m = self._metadata.get(source_id, {}).get(final_signature)
        return json.loads(json.dumps(m)) if m is not None else None

    # ------------------------------------------------------------------ lazy sources
    def register_tiled_source(self, image: TiledPipelineImage, *, band_rows: int = 2048) -> str:
        """Source id of a lazy handle WITHOUT densifying it on the host: the digest of ``register_source``
        (sha256 over str(shape), str(dtype) and the row-major bytes) is fed band by band while the same
        bands are uploaded; the frame becomes the resident source."""
        import torch

        from . import ingest
        from .mosaic import RowSource

        be = self._executor.backend
        src = RowSource(image)
        h, w = src.shape
        digest = hashlib.sha256()
        digest.update(str((h, w)).encode("utf-8"))
        digest.update(str(src.dtype).encode("utf-8"))
        parts = []
        for r0 in range(0, h, band_rows):
            r1 = min(h, r0 + band_rows)
            band = np.ascontiguousarray(src[r0:r1])
            digest.update(band.tobytes())
            parts.append(ingest.upload_rows(be, band, 0, r1 - r0))
        sid = digest.hexdigest()
        self._put(sid, sid, parts[0] if len(parts) == 1 else torch.cat(parts, dim=0))
        with self._lock:
            self._metadata.setdefault(sid, {})[sid] = {"version": 1, "source_id": sid, "final_signature": sid, "steps": []}
        return sid

    # ------------------------------------------------------------------ disk tier (reference layout)
    def _write_disk(self, source_id: str, signature: str, array: np.ndarray,
                    tiles: Optional[List[Tuple[TileBox, np.ndarray]]], tile_size) -> None:
        d = self._cache_directory
        if d is None:
            return
        if tiles is None:
            path = d / f"{source_id}_{signature}.npy"
            tmp = path.with_suffix(".npy.tmp")
            payload = lambda fh: np.save(fh, array)
        else:
            path = d / f"{source_id}_{signature}.npz"
            tmp = path.with_suffix(".npz.tmp")
            meta = {"type": "tiles", "shape": list(array.shape), "dtype": str(array.dtype),
                    "tile_size": list(tile_size) if tile_size is not None else None,
                    "boxes": [list(box) for box, _ in tiles]}
            arrays = {f"tile_{i}": t for i, (_, t) in enumerate(tiles)}
            arrays["metadata"] = np.array(json.dumps(meta))
            payload = lambda fh: np.savez(fh, **arrays)
        try:
            with tmp.open("wb") as fh:
                payload(fh)
                fh.flush()
                os.fsync(fh.fileno())
            os.replace(tmp, path)
        except OSError:
            with contextlib.suppress(FileNotFoundError):
                tmp.unlink()
            raise

    def get_cached_image(self, source_id: str, signature: str) -> Optional[np.ndarray]:
        """The image stored under ``signature``: from HBM if resident, else from the disk tier."""
        t = self._get(source_id, signature)
        if t is not None:
            return self._executor.backend.to_host(t)
        d = self._cache_directory
        if d is None:
            return None
        dense = d / f"{source_id}_{signature}.npy"
        if dense.exists():
            return np.load(dense, allow_pickle=False)
        tiled = d / f"{source_id}_{signature}.npz"
        if tiled.exists():
            with np.load(tiled, allow_pickle=False) as z:
                meta = json.loads(str(z["metadata"]))
                out = np.zeros(tuple(meta["shape"]), dtype=np.dtype(meta["dtype"]))
                for i, (left, top, right, bottom) in enumerate(meta["boxes"]):
                    out[top:bottom, left:right, ...] = z[f"tile_{i}"]
            return out
        return None

    # ------------------------------------------------------------------ compute
    def _stream_tiles(self, tensor, tile_size, emit) -> Tuple[np.ndarray, List[Tuple[TileBox, np.ndarray]]]:
        """Download ``tensor`` band by band (one band = one row of tiles) and hand out its tiles in
        row-major order; the copy of band k + 1 is in flight while band k is being emitted."""
        import torch

        be = self._executor.backend
        h, w = int(tensor.shape[0]), int(tensor.shape[1])
        tw, th = tile_size if tile_size is not None else (w, h)
        out = be.pinned_empty(tuple(tensor.shape), np.dtype(str(tensor.dtype).replace("torch.", "")))
        host = torch.from_numpy(out)
        copy_stream = torch.cuda.Stream(device=be.device)
        copy_stream.wait_stream(torch.cuda.current_stream(be.device))
        bands = []
        for top in range(0, h, th):
            bottom = min(h, top + th)
            with torch.cuda.stream(copy_stream):
                host[top:bottom].copy_(tensor[top:bottom], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            bands.append((top, bottom, ev))
        tiles: List[Tuple[TileBox, np.ndarray]] = []
        boxes = list(iter_tile_boxes(w, h, (tw, th)))
        bi = 0
        for top, bottom, ev in bands:
            ev.synchronize()
            while bi < len(boxes) and boxes[bi][1] == top:
                left, t0, right, b0 = boxes[bi]
                tile = np.array(out[t0:b0, left:right, ...], copy=True)
                tiles.append((boxes[bi], tile))
                emit(boxes[bi], tile)
                bi += 1
        torch.cuda.current_stream(be.device).wait_stream(copy_stream)
        return out, tiles

    def compute(self, source_id: str, image, steps: Sequence[Any], *,
                cancel_event: Optional[threading.Event] = None,
                progress: Optional[Callable[[int], None]] = None,
                incremental: Optional[Callable[[PipelineCacheTileUpdate], None]] = None,
                tile_size: Optional[Tuple[int, int]] = None) -> DeviceCacheResult:
        """Evaluate ``steps``; ``image`` (ndarray or ``TiledPipelineImage``) is only uploaded when the
        source is not resident any more.  ``incremental`` receives one ``PipelineCacheTileUpdate`` per
        tile of the final image (``tile_size`` = (w, h); default: the handle's hint, else one tile)."""
        be = self._executor.backend
        final_signature, records = self.predict(source_id, steps)
        if tile_size is None and isinstance(image, TiledPipelineImage):
            tile_size = image.tile_size
        current = self._get(source_id, source_id)
        if current is None:
            if image is None:
                raise KeyError(f"source {source_id[:12]}… is not resident and no image was given")
            if isinstance(image, TiledPipelineImage):
                if self.register_tiled_source(image) != source_id:
                    raise ValueError("image does not match source_id")
                current = self._get(source_id, source_id)
            else:
                if cache_keys.source_id(image) != source_id:
                    raise ValueError("image does not match source_id")
                current = be.to_device(np.ascontiguousarray(image))
                self._put(source_id, source_id, current)
        total = max(1, len(steps))
        ran: List[int] = []
        for index, (step, record) in enumerate(zip(steps, records)):
            if cancel_event is not None and cancel_event.is_set():
                raise OperationCancelled()
            cached = self._get(source_id, record.signature)
            if cached is not None:
                current = cached
            else:
                if step.enabled:
                    current = self._executor.run_on_device(step.name, current, step.params)
                    ran.append(index)
                # a disabled step maps its signature to the unchanged tensor (no copy: tensors are never mutated)
                self._put(source_id, record.signature, current)
            if progress is not None:
                progress(int(((index + 1) / total) * 100))
        meta = {"version": 1, "source_id": source_id, "final_signature": final_signature,
                "steps": [r.to_dict() for r in records]}
        with self._lock:
            self._metadata.setdefault(source_id, {})[final_signature] = meta
        tiles = None
        if incremental is not None and current.dim() >= 2:
            shape, dtype = tuple(int(v) for v in current.shape), np.dtype(str(current.dtype).replace("torch.", ""))
            step_sig = records[-1].signature if records else final_signature

            def emit(box, tile):
                if cancel_event is not None and cancel_event.is_set():
                    raise OperationCancelled()
                incremental(PipelineCacheTileUpdate(source_id, final_signature, step_sig, total, total, box, tile, shape, dtype,
                                                    tile_size, from_cache=not ran))

            host, tiles = self._stream_tiles(current, tile_size, emit)
        else:
            host = be.to_host(current)
        self._write_disk(source_id, final_signature, host, tiles if tile_size is not None else None, tile_size)
        return DeviceCacheResult(source_id, final_signature, host, records,
                                 json.loads(json.dumps(meta)), current, tuple(ran))


__all__ = ["DeviceCacheResult", "DevicePipelineCache", "OperationCancelled", "PipelineCacheTileUpdate"]
