"""``pipeline_cache`` key contract (reference: ``processing/pipeline_cache.py``).

The reference memoises per-step outputs under a signature chain that depends only on
``(previous signature, step.name, step.enabled, step.params)`` (``predict:291-313``,
``_hash_payload:52-57``, ``_normalise_value:40-49``) seeded by
``source_id = sha256(str(shape) + str(dtype) + bytes)`` (``register_source:256-264``).  A GPU-backed
step with the same name and params therefore hits the same cache entries; these helpers restate
the key derivation so that can be tested without the Qt-dependent cache module.
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import dataclass
from typing import Any, Dict, List, Mapping, Sequence, Tuple

import numpy as np


def normalise_value(value: Any) -> Any:
    if value is None or isinstance(value, (str, int, float, bool)):
        return value
    if isinstance(value, (list, tuple, set)):
        return [normalise_value(v) for v in value]
    if isinstance(value, Mapping):
        return {k: normalise_value(value[k]) for k in sorted(value)}
    return repr(value)


def hash_payload(payload: Mapping[str, Any]) -> str:
    blob = json.dumps(payload, sort_keys=True, separators=(",", ":")).encode("utf-8")
    return hashlib.sha256(blob).hexdigest()


def source_id(image: np.ndarray) -> str:
    a = np.ascontiguousarray(image)
    h = hashlib.sha256()
    h.update(str(a.shape).encode("utf-8"))
    h.update(str(a.dtype).encode("utf-8"))
    h.update(a.tobytes())
    return h.hexdigest()


@dataclass(frozen=True)
class StepRecord:
    name: str
    enabled: bool
    params: Dict[str, Any]
    signature: str
    index: int

    def to_dict(self) -> Dict[str, Any]:
        """Same record the reference persists (processing/pipeline_cache.py:70-77)."""
        return {"name": self.name, "enabled": self.enabled,
                "params": {k: normalise_value(v) for k, v in self.params.items()},
                "signature": self.signature, "index": self.index}


def predict(source: str, steps: Sequence[Any]) -> Tuple[str, List[StepRecord]]:
    """Final signature and per-step records for ``steps`` applied to ``source``."""
    sig = source
    out: List[StepRecord] = []
    for i, s in enumerate(steps):
        sig = hash_payload({"previous": sig, "name": s.name, "enabled": bool(s.enabled),
                            "params": normalise_value(s.params)})
        out.append(StepRecord(s.name, bool(s.enabled), dict(s.params), sig, i))
    return sig, out


def dense_cache_filename(source: str, signature: str) -> str:
    """``{source_id}_{signature}.npy`` (processing/pipeline_cache.py:721-799)."""
    return f"{source}_{signature}.npy"


__all__ = ["StepRecord", "dense_cache_filename", "hash_payload", "normalise_value", "predict", "source_id"]
