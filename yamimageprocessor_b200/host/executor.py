"""``GpuExecutor`` implementation for the reference's ``PipelineManager``.

``B200Executor.execute(step, image)`` satisfies the protocol at
``processing/pipeline_manager.py:69-73`` (dispatch site ``:448-454``): it receives the dense
array (the whole N-D stack for time-lapse input), runs the step named ``step.name`` on the GPU
and returns a fresh C-contiguous host array.  ``execute_chain`` runs several consecutive steps
with one upload and one download (used by this package's ``PipelineManager`` mirror) and keeps a
binary segmentation run (``Adaptive`` -> rectangular morphology -> ``ConnectedComponents``) packed
1 bit/pixel between its kernels -- same result, the 8-bit mask never reaches HBM.

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
    def __init__(self, backend: Optional[Backend] = None, device: Optional[int] = None) -> None:
        self._backend = backend
        self._device = device
        self.calls: list[str] = []  # step names executed, in order (diagnostics / tests)

    @property
    def backend(self) -> Backend:
        if self._backend is None:
            device = self._device
            if device is None:
                # one process per GPU: the device this process selected (torch.cuda.set_device(LOCAL_RANK))
                import torch

                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            self._backend = get_backend(int(device))
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
        return be.to_host(self.run_chain_on_device(steps, be.to_device(np.asarray(image))))

    # ------------------------------------------------------------------ binary-run fusion
    _MORPH_OPS = {"Opening": 2, "Closing": 3, "Dilation": 1, "Erosion": 0}
    _BIT_BLOCKS = (3, 5, 7, 11, 15)

    def _bit_run_length(self, steps: Sequence[Any], i: int, tensor) -> int:
        """Length of the run starting at steps[i] that can stay 1 bit/pixel: ``Adaptive`` followed
        by rectangular morphology steps and, optionally, ``ConnectedComponents`` (0: no such run).
        A thresholded mask is binary, so the run gives the same result as the step-by-step chain."""
        import torch

        step = steps[i]
        if step.name != "Adaptive" or int(step.params.get("block_size", 11)) not in self._BIT_BLOCKS:
            return 0
        plane = tensor.dim() == 2 or (tensor.dim() == 3 and tensor.shape[-1] not in (3, 4))
        if not plane or tensor.dtype not in (torch.uint8, torch.uint16):
            return 0
        j = i + 1
        while j < len(steps) and steps[j].name in self._MORPH_OPS:
            p = steps[j].params
            k = int(p.get("kernel_size", 3))
            if str(p.get("kernel_shape", "Rectangular")).lower() not in ("rectangular", "rect") or not (1 <= k <= 31):
                break
            j += 1
        if j < len(steps) and steps[j].name == "ConnectedComponents":
            j += 1
        return j - i if j - i >= 2 else 0

    def _run_bit_run(self, run: Sequence[Any], tensor):
        be = self.backend
        first = run[0].params
        width = int(tensor.shape[-1])
        self.calls.append("Adaptive")
        bits = be.adaptive_threshold_bits(tensor, int(first.get("block_size", 11)), float(first.get("C", 2)))
        rest = list(run[1:])
        i = 0
        while i < len(rest):
            step = rest[i]
            self.calls.append(step.name)
            if step.name == "ConnectedComponents":
                return be.ccl_label_bits(bits, width)[0]
            k, it = int(step.params.get("kernel_size", 3)), int(step.params.get("iterations", 1))
            nxt = rest[i + 1] if i + 1 < len(rest) else None
            if (step.name == "Opening" and nxt is not None and nxt.name == "Closing"
                    and int(nxt.params.get("kernel_size", 3)) == k and int(nxt.params.get("iterations", 1)) == it):
                # open then close with the same element: one launch (erode, dilate x2, erode in registers)
                self.calls.append(nxt.name)
                bits = be.bits_morph(bits, width, 4, k, it)
                i += 2
                continue
            bits = be.bits_morph(bits, width, self._MORPH_OPS[step.name], k, it)
            i += 1
        return be.bits_unpack(bits, width)

    def run_chain_on_device(self, steps: Sequence[Any], tensor):
        """Consecutive steps on a device tensor; intermediates never leave HBM and a binary
        segmentation run (see ``_bit_run_length``) keeps its mask packed 1 bit/pixel."""
        active = [s for s in steps if getattr(s, "enabled", True)]
        i = 0
        while i < len(active):
            n = self._bit_run_length(active, i, tensor)
            if n:
                tensor = self._run_bit_run(active[i:i + n], tensor)
                i += n
            else:
                tensor = self.run_on_device(active[i].name, tensor, active[i].params)
                i += 1
        return tensor


__all__ = ["B200Executor"]
