"""``GpuExecutor`` implementation for the reference's ``PipelineManager``.

``B200Executor.execute(step, image)`` satisfies the protocol at
``processing/pipeline_manager.py:69-73`` (dispatch site ``:448-454``): it receives the dense
array (the whole N-D stack for time-lapse input), runs the step named ``step.name`` on the GPU
and returns a fresh C-contiguous host array.  ``execute_chain`` runs several consecutive steps
with one upload and one download (used by this package's ``PipelineManager`` mirror).

Steps are looked up by NAME in ``steps.DEVICE_STEPS`` — the same names/params that feed the
``pipeline_cache`` signatures — so a pipeline built by the reference's own builders dispatches
here unchanged.  Unknown names raise ``KeyError``; nothing ever falls back to ``step.function``.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional, Sequence

import numpy as np

from ..backend import Backend, get_backend
from .steps import DEVICE_STEPS


class B200Executor:
    def __init__(self, backend: Optional[Backend] = None, device: int = 0) -> None:
        self._backend = backend
        self._device = device
        self.calls: list[str] = []  # step names executed, in order (diagnostics / tests)

    @property
    def backend(self) -> Backend:
        if self._backend is None:
            self._backend = get_backend(self._device)
        return self._backend

    @staticmethod
    def supports(step_name: str) -> bool:
        return step_name in DEVICE_STEPS

    def _lookup(self, name: str):
        try:
            return DEVICE_STEPS[name]
        except KeyError:
            raise KeyError(
                f"step '{name}' has no B200 kernel (known: {sorted(DEVICE_STEPS)}); "
                "this backend has no CPU fallback"
            ) from None

    def run_on_device(self, name: str, tensor, params: Mapping[str, Any]):
        self.calls.append(name)
        return self._lookup(name)(self.backend, tensor, params)

    def execute(self, step, image: np.ndarray) -> np.ndarray:
        be = self.backend
        out = self.run_on_device(step.name, be.to_device(np.asarray(image)), step.params)
        return be.to_host(out)

    def execute_chain(self, steps: Sequence[Any], image: np.ndarray) -> np.ndarray:
        be = self.backend
        t = be.to_device(np.asarray(image))
        for step in steps:
            if getattr(step, "enabled", True):
                t = self.run_on_device(step.name, t, step.params)
        return be.to_host(t)


__all__ = ["B200Executor"]
