"""Host-side mirror of the reference's pipeline runtime for the hot path.

Same public names, argument meaning and error behaviour as
``processing/pipeline_manager.py`` in the reference (``StepExecutionMetadata:45``,
``GpuExecutor:69``, ``PipelineStep:80``, ``PipelineState:173``,
``PipelineManager:189``), written from scratch so the parity tests of this
repository read like the reference's own (``tests/test_pipeline_manager.py``,
``tests/test_processing_pipeline_manager_gpu.py``) and run where the reference
checkout is absent.  Inside the real application the reference's own manager is
used and only the plugin / executor of this package are loaded.

One extension, invisible to reference callers: when the configured executor
offers ``execute_chain`` (``B200Executor`` does), ``apply`` hands every maximal
run of consecutive GPU steps to it in one call so intermediates stay in HBM.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Protocol, Sequence, Tuple, Union

import numpy as np

from .tiles import TileBox, TiledPipelineImage

LOGGER = logging.getLogger(__name__)

PipelineImage = Union[np.ndarray, TiledPipelineImage]
PipelineChangeListener = Callable[[str, Dict[str, Any]], None]


def _looks_like_colour(a: np.ndarray) -> bool:
    return a.ndim == 3 and a.shape[2] in (3, 4)


@dataclass
class StepExecutionMetadata:
    """Execution hints of a step (reference: processing/pipeline_manager.py:45-66)."""

    supports_inplace: bool = False
    requires_gpu: bool = False

    def to_dict(self) -> Dict[str, Any]:
        return {"supports_inplace": self.supports_inplace, "requires_gpu": self.requires_gpu}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "StepExecutionMetadata":
        return cls(bool(data.get("supports_inplace", False)), bool(data.get("requires_gpu", False)))

    def is_default(self) -> bool:
        return not self.supports_inplace and not self.requires_gpu


class GpuExecutor(Protocol):
    """``execute(step, image) -> ndarray | None`` (reference: processing/pipeline_manager.py:69-73)."""

    def execute(self, step: "PipelineStep", image: np.ndarray) -> np.ndarray: ...


@dataclass
class PipelineStep:
    """One named, parameterised step (reference: processing/pipeline_manager.py:80-170)."""

    name: str
    function: Callable[..., PipelineImage]
    enabled: bool = True
    params: Dict[str, Any] = field(default_factory=dict)
    execution: StepExecutionMetadata = field(default_factory=StepExecutionMetadata)
    supports_tiled_input: bool = False
    stage: Optional[Any] = field(default=None, repr=False, compare=False)

    def apply(self, image: PipelineImage) -> PipelineImage:
        if not self.enabled:
            return image
        operand: PipelineImage = image
        if isinstance(operand, TiledPipelineImage) and not self.supports_tiled_input:
            operand = operand.to_array()
        produced = self.function(operand, **self.params)
        if produced is None:
            produced = operand
        inplace_ok = (
            self.execution.supports_inplace
            and isinstance(operand, np.ndarray)
            and isinstance(produced, np.ndarray)
        )
        if inplace_ok:
            if produced is operand:
                return operand
            if produced.shape == operand.shape and produced.dtype == operand.dtype:
                operand[...] = produced
                return operand
        return produced

    def clone(self) -> "PipelineStep":
        return PipelineStep(
            self.name,
            self.function,
            self.enabled,
            dict(self.params),
            StepExecutionMetadata(self.execution.supports_inplace, self.execution.requires_gpu),
            self.supports_tiled_input,
            self.stage,
        )

    def to_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {"name": self.name, "enabled": self.enabled, "params": dict(self.params)}
        if not self.execution.is_default():
            out["execution"] = self.execution.to_dict()
        if self.supports_tiled_input:
            out["supports_tiled_input"] = True
        if self.stage is not None:
            value = getattr(self.stage, "value", None)
            out["stage"] = value if value is not None else str(self.stage)
        return out

    @classmethod
    def from_dict(cls, data: Dict[str, Any], function: Callable[..., PipelineImage]) -> "PipelineStep":
        stage = data.get("stage")
        if isinstance(stage, str):
            try:
                from .plugin import ModuleStage

                stage = ModuleStage(stage)
            except Exception:
                stage = None
        return cls(
            name=data["name"],
            function=function,
            enabled=bool(data.get("enabled", True)),
            params=dict(data.get("params", {})),
            execution=StepExecutionMetadata.from_dict(data.get("execution", {})),
            supports_tiled_input=bool(data.get("supports_tiled_input", False)),
            stage=stage,
        )


@dataclass
class PipelineState:
    """History snapshot (reference: processing/pipeline_manager.py:173-186)."""

    steps: List[PipelineStep]
    image: Optional[np.ndarray] = None
    cache_signature: Optional[str] = None

    def clone(self) -> "PipelineState":
        img = None if self.image is None else self.image.copy()
        return PipelineState([s.clone() for s in self.steps], img, self.cache_signature)


def _as_dir(path: Optional[Union[str, os.PathLike]]) -> Optional[Path]:
    if path is None:
        return None
    p = Path(path)
    p.mkdir(parents=True, exist_ok=True)
    return p


class PipelineManager:
    """Ordered, editable list of steps with undo/redo and dense / stack / tiled execution."""

    _DEFAULT_CACHE_DIR: Optional[Path] = None
    _DEFAULT_RECOVERY_ROOT: Optional[Path] = None

    def __init__(
        self,
        steps: Optional[Iterable[PipelineStep]] = None,
        *,
        cache_dir: Optional[Union[str, os.PathLike]] = None,
        recovery_root: Optional[Union[str, os.PathLike]] = None,
        gpu_executor: Optional[GpuExecutor] = None,
    ) -> None:
        self._template: List[PipelineStep] = [s.clone() for s in (steps or [])]
        self._steps: List[PipelineStep] = [s.clone() for s in self._template]
        self._undo: List[PipelineState] = []
        self._redo: List[PipelineState] = []
        self._listeners: List[PipelineChangeListener] = []
        self._gpu_executor = gpu_executor
        self._cache_directory = _as_dir(cache_dir if cache_dir is not None else self._DEFAULT_CACHE_DIR)
        self._recovery_root = _as_dir(recovery_root if recovery_root is not None else self._DEFAULT_RECOVERY_ROOT)

    # ---- configuration ---------------------------------------------------------------------
    @classmethod
    def set_default_cache_directory(cls, path) -> None:
        cls._DEFAULT_CACHE_DIR = _as_dir(path)

    @classmethod
    def set_default_recovery_root(cls, path) -> None:
        cls._DEFAULT_RECOVERY_ROOT = _as_dir(path)

    @property
    def cache_directory(self) -> Optional[Path]:
        return self._cache_directory

    @property
    def recovery_root(self) -> Optional[Path]:
        return self._recovery_root

    def set_cache_directory(self, path) -> None:
        self._cache_directory = _as_dir(path)

    def set_recovery_root(self, path) -> None:
        self._recovery_root = _as_dir(path)

    def set_gpu_executor(self, executor: Optional[GpuExecutor]) -> None:
        self._gpu_executor = executor

    # ---- introspection ---------------------------------------------------------------------
    def __iter__(self) -> Iterator[PipelineStep]:
        return iter(self._steps)

    @property
    def steps(self) -> Tuple[PipelineStep, ...]:
        return tuple(self._steps)

    def iter_enabled_steps(self) -> Iterator[PipelineStep]:
        return (s for s in self._steps if s.enabled)

    def template_steps(self) -> Tuple[PipelineStep, ...]:
        return tuple(s.clone() for s in self._template)

    def get_step(self, identifier: Union[int, str]) -> PipelineStep:
        if isinstance(identifier, int):
            return self._steps[identifier]
        for s in self._steps:
            if s.name == identifier:
                return s
        raise KeyError(f"No pipeline step named '{identifier}'")

    def to_dict(self) -> Dict[str, Any]:
        return {"steps": [s.to_dict() for s in self._steps]}

    def clone(self) -> "PipelineManager":
        twin = PipelineManager(
            self._template,
            cache_dir=self._cache_directory,
            recovery_root=self._recovery_root,
            gpu_executor=self._gpu_executor,
        )
        twin._steps = [s.clone() for s in self._steps]
        return twin

    # ---- editing ---------------------------------------------------------------------------
    def reset(self) -> None:
        self._steps = [s.clone() for s in self._template]
        self.clear_history()
        self._emit("pipeline_reset", steps=tuple(self._steps))

    def replace_steps(self, steps: Iterable[PipelineStep], *, update_template: bool = False,
                      preserve_history: bool = False) -> None:
        self._steps = [s.clone() for s in steps]
        if update_template:
            self._template = [s.clone() for s in self._steps]
        if not preserve_history:
            self.clear_history()
        self._emit("steps_replaced", steps=tuple(self._steps))

    def add_step(self, step: PipelineStep, index: Optional[int] = None) -> None:
        at = len(self._steps) if index is None else index
        self._steps.insert(at, step)
        self._emit("step_added", step=step, index=at)

    def remove_step(self, index: int) -> PipelineStep:
        gone = self._steps.pop(index)
        self._emit("step_removed", step=gone, index=index)
        return gone

    def move_step(self, old_index: int, new_index: int) -> None:
        s = self._steps.pop(old_index)
        self._steps.insert(new_index, s)
        self._emit("steps_reordered", step=s, old_index=old_index, new_index=new_index, steps=tuple(self._steps))

    def swap_steps(self, index_a: int, index_b: int) -> None:
        st = self._steps
        st[index_a], st[index_b] = st[index_b], st[index_a]
        self._emit("steps_swapped", first_index=index_a, second_index=index_b, steps=tuple(st))

    def set_order(self, order: Iterable[str]) -> None:
        """Named steps first, in the given order; the rest keep their relative order."""
        by_name = {s.name: s for s in self._steps}
        picked: List[PipelineStep] = []
        for name in order:
            if name not in by_name:
                raise KeyError(f"Unknown pipeline step '{name}'")
            picked.append(by_name.pop(name))
        rest = [s for s in self._steps if s.name in by_name]
        self._steps = picked + rest
        self._emit("steps_reordered", steps=tuple(self._steps))

    def set_step_enabled(self, identifier: Union[int, str], enabled: bool) -> None:
        s = self.get_step(identifier)
        if s.enabled != enabled:
            s.enabled = enabled
            self._emit("step_state_changed", step=s, enabled=enabled)

    def toggle_step(self, identifier: Union[int, str]) -> bool:
        s = self.get_step(identifier)
        s.enabled = not s.enabled
        self._emit("step_state_changed", step=s, enabled=s.enabled)
        return s.enabled

    def update_step_params(self, identifier: Union[int, str], params: Dict[str, Any], *, replace: bool = False) -> None:
        s = self.get_step(identifier)
        if replace:
            s.params = dict(params)
        else:
            s.params.update(params)
        self._emit("step_params_updated", step=s, replace=replace)

    # ---- execution -------------------------------------------------------------------------
    def apply(self, image: PipelineImage) -> PipelineImage:
        if isinstance(image, TiledPipelineImage):
            return self._apply_tiled(image)
        active = list(self.iter_enabled_steps())
        current: PipelineImage = image
        # The reference copies the input up front so that in-place CPU steps cannot touch the caller's
        # array (processing/pipeline_manager.py:400).  A step routed to the GPU executor never mutates
        # its input and always returns a fresh array, so the (full-frame) copy is skipped when the
        # first enabled step goes to the executor; every other case keeps the defensive copy.
        first_on_gpu = bool(active) and active[0].execution.requires_gpu and self._gpu_executor is not None
        if isinstance(image, np.ndarray) and not first_on_gpu:
            current = image.copy()
        out = self._run_steps(active, current)
        if out is image and isinstance(image, np.ndarray):
            out = image.copy()  # nothing ran (or the executor returned None): still hand back a copy
        return out

    def _run_steps(self, steps: Sequence[PipelineStep], current: PipelineImage) -> PipelineImage:
        chain = getattr(self._gpu_executor, "execute_chain", None)
        i = 0
        while i < len(steps):
            step = steps[i]
            if chain is not None and step.execution.requires_gpu:
                j = i
                while j < len(steps) and steps[j].execution.requires_gpu:
                    j += 1
                if j - i > 1:
                    dense = current if isinstance(current, np.ndarray) else current.to_array()
                    out = chain(steps[i:j], dense)
                    current = dense if out is None else out
                    i = j
                    continue
            current = self._run_step(step, current)
            i += 1
        return current

    def _apply_tiled(self, image: TiledPipelineImage) -> PipelineImage:
        active = list(self.iter_enabled_steps())
        if not active:
            return image
        if any(s.supports_tiled_input for s in active):
            # steps that understand the lazy handle pull their own tiles (sharded GPU steps do)
            return self._run_steps(active, image)
        canvas: Optional[np.ndarray] = None
        shape = image.infer_shape()
        for box, tile in image.iter_tiles(image.tile_size):
            piece: PipelineImage = np.array(tile, copy=True)
            for s in active:
                piece = self._run_step(s, piece)
                if isinstance(piece, TiledPipelineImage):
                    piece = piece.to_array()
            arr = np.asarray(piece)
            if canvas is None:
                canvas = np.zeros(shape, dtype=arr.dtype)
            self._paste_tile(canvas, box, arr)
        return image.to_array() if canvas is None else canvas

    @staticmethod
    def _paste_tile(target: np.ndarray, box: TileBox, tile: np.ndarray) -> None:
        left, top, right, bottom = box
        target[top:bottom, left:right, ...] = tile

    def _run_step(self, step: PipelineStep, image: PipelineImage) -> PipelineImage:
        if step.execution.requires_gpu:
            dense = image if isinstance(image, np.ndarray) else image.to_array()
            if self._gpu_executor is not None:
                out = self._gpu_executor.execute(step, dense)
                return dense if out is None else out
            LOGGER.warning(
                "Step '%s' requires GPU execution but no executor is configured; falling back to CPU.",
                step.name,
            )
            return step.apply(dense)
        if isinstance(image, np.ndarray) and self._requires_slice_processing(image):
            return self._apply_slice_wise(step, image)
        return step.apply(image)

    @staticmethod
    def _requires_slice_processing(array: np.ndarray) -> bool:
        if array.ndim <= 2:
            return False
        return not (array.ndim == 3 and _looks_like_colour(array))

    def _apply_slice_wise(self, step: PipelineStep, array: np.ndarray) -> np.ndarray:
        planes: List[np.ndarray] = []
        for plane in array:
            out = step.apply(plane)
            planes.append(out.to_array() if isinstance(out, TiledPipelineImage) else np.asarray(out))
        if not planes:
            return array.copy()
        try:
            return np.stack(planes, axis=0)
        except ValueError:
            return np.array(planes, dtype=object)

    @staticmethod
    def extract_preview(array: np.ndarray, axis: int = 0) -> np.ndarray:
        if array.ndim <= 2 or (array.ndim == 3 and _looks_like_colour(array)):
            return np.asarray(array)
        axis = min(max(axis, 0), array.ndim - 1)
        return np.take(array, array.shape[axis] // 2, axis=axis)

    # ---- history ---------------------------------------------------------------------------
    def _snapshot(self, image, signature) -> PipelineState:
        return PipelineState([s.clone() for s in self._steps], None if image is None else image.copy(), signature)

    def push_state(self, *, image: Optional[np.ndarray] = None, cache_signature: Optional[str] = None) -> None:
        self._undo.append(self._snapshot(image, cache_signature))
        self._redo.clear()

    def _restore(self, source: List[PipelineState], sink: List[PipelineState], tag: str, image, signature):
        if not source:
            return None
        sink.append(self._snapshot(image, signature))
        state = source.pop()
        self._steps = [s.clone() for s in state.steps]
        self._emit("pipeline_restored", source=tag, steps=tuple(self._steps))
        return state.clone()

    def undo(self, *, current_image: Optional[np.ndarray] = None, current_cache_signature: Optional[str] = None):
        return self._restore(self._undo, self._redo, "undo", current_image, current_cache_signature)

    def redo(self, *, current_image: Optional[np.ndarray] = None, current_cache_signature: Optional[str] = None):
        return self._restore(self._redo, self._undo, "redo", current_image, current_cache_signature)

    def clear_history(self) -> None:
        self._undo.clear()
        self._redo.clear()

    def history_depth(self) -> Tuple[int, int]:
        return len(self._undo), len(self._redo)

    def can_undo(self) -> bool:
        return bool(self._undo)

    def can_redo(self) -> bool:
        return bool(self._redo)

    # ---- listeners -------------------------------------------------------------------------
    def add_change_listener(self, listener: PipelineChangeListener) -> None:
        if listener not in self._listeners:
            self._listeners.append(listener)

    def remove_change_listener(self, listener: PipelineChangeListener) -> None:
        if listener in self._listeners:
            self._listeners.remove(listener)

    def _emit(self, event: str, **meta: Any) -> None:
        for listener in tuple(self._listeners):
            try:
                listener(event, dict(meta))
            except Exception:  # listeners must not break the pipeline
                LOGGER.debug("Pipeline change listener failed", exc_info=True)


__all__ = [
    "GpuExecutor",
    "PipelineChangeListener",
    "PipelineImage",
    "PipelineManager",
    "PipelineState",
    "PipelineStep",
    "StepExecutionMetadata",
]
