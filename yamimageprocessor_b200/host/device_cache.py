"""Device-resident tier of the reference's ``PipelineCache``.

``processing/pipeline_cache.py`` memoises every step output under a signature chain
(``predict:291-313``) and, per step, copies the array two or three times, ``np.save``s it and
fsyncs (``_compute_dense:352-394``).  This class keeps the same keys and the same
``register_source / predict / compute / discard_cache`` surface but stores the per-step outputs as
CUDA tensors: a re-run after a parameter change restarts from the last step whose signature still
matches, entirely on the device, and only the final image is downloaded.  Entries are evicted
least-recently-used beyond ``max_bytes``.  Steps run through ``B200Executor`` by NAME (the names
that feed the signatures); unknown names raise ``KeyError`` — there is no CPU fallback.
"""
from __future__ import annotations

import json
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import cache_keys
from .cache_keys import StepRecord
from .executor import B200Executor


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
    def __init__(self, executor: Optional[B200Executor] = None, max_bytes: int = 16 << 30) -> None:
        self._executor = executor or B200Executor()
        self._max_bytes = int(max_bytes)
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

    def compute(self, source_id: str, image: Optional[np.ndarray], steps: Sequence[Any], *,
                cancel_event: Optional[threading.Event] = None,
                progress: Optional[Callable[[int], None]] = None) -> DeviceCacheResult:
        """Evaluate ``steps``; ``image`` is only uploaded when the source is not resident any more."""
        be = self._executor.backend
        final_signature, records = self.predict(source_id, steps)
        current = self._get(source_id, source_id)
        if current is None:
            if image is None:
                raise KeyError(f"source {source_id[:12]}… is not resident and no image was given")
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
        return DeviceCacheResult(source_id, final_signature, be.to_host(current), records,
                                 json.loads(json.dumps(meta)), current, tuple(ran))


__all__ = ["DeviceCacheResult", "DevicePipelineCache", "OperationCancelled"]
