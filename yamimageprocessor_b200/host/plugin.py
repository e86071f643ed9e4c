"""Plugin contract mirror: ``ModuleBase`` / ``ModuleStage`` / ``ModuleMetadata``.

Reference: ``plugins/module_base.py`` (``ModuleStage:19``, ``ModuleMetadata:27``,
``MenuEntry:41``, ``ModuleBase:50`` with ``create_pipeline_step:133``,
``process:146``, ``sanitize_parameters:113``, ``pipeline_execution_metadata:123``,
``supports_tiled_input:128``).  ``binding()`` returns the reference's own classes
when the application is importable (so ``AppCore.register_module`` accepts the
subclasses: it checks ``issubclass(cls, ModuleBase)``, ``core/app_core.py:753-757``)
and this mirror otherwise.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Mapping, Sequence, Tuple

from .pipeline import PipelineStep, StepExecutionMetadata


class ModuleStage(Enum):
    PREPROCESSING = "preprocessing"
    SEGMENTATION = "segmentation"
    ANALYSIS = "analysis"


@dataclass(frozen=True)
class ModuleMetadata:
    identifier: str
    title: str
    stage: ModuleStage
    description: str = ""
    menu_path: Tuple[str, ...] = ("Pre-Processing",)
    shortcut: str | None = None
    default_enabled: bool = False


@dataclass(frozen=True)
class MenuEntry:
    path: Tuple[str, ...]
    text: str
    description: str = ""
    shortcut: str | None = None


class ModuleBase(ABC):
    """Base class of processing modules; subclasses implement ``_build_metadata`` and ``process``."""

    def __init__(self) -> None:
        self._metadata = self._build_metadata()

    @property
    def metadata(self) -> ModuleMetadata:
        return self._metadata

    @abstractmethod
    def _build_metadata(self) -> ModuleMetadata: ...

    @abstractmethod
    def process(self, image, **kwargs: Any): ...

    # parameter registry: the reference pulls these from ui/control_metadata.py; modules of this
    # package carry their own table (same defaults / ranges / coercions) so no UI import is needed
    def parameter_metadata(self) -> Mapping[str, Any]:
        return {}

    def default_parameters(self) -> Dict[str, Any]:
        return {k: m.default for k, m in self.parameter_metadata().items() if m.default is not None}

    def sanitize_parameters(self, params: Mapping[str, Any]) -> Dict[str, Any]:
        merged: Dict[str, Any] = dict(self.default_parameters())
        merged.update(params)
        for key, meta in self.parameter_metadata().items():
            if key in merged:
                merged[key] = meta.coerce(merged[key])
        return merged

    def menu_entries(self) -> Sequence[MenuEntry]:
        m = self.metadata
        return (MenuEntry(m.menu_path, m.title, m.description, m.shortcut),)

    def activate(self, pane) -> None:
        raise NotImplementedError(f"{type(self).__name__} does not implement an activation handler")

    def pipeline_execution_metadata(self) -> StepExecutionMetadata:
        return StepExecutionMetadata()

    def supports_tiled_input(self) -> bool:
        return False

    def create_pipeline_step(self) -> PipelineStep:
        m = self.metadata
        return PipelineStep(
            name=m.identifier,
            function=self.process,
            enabled=m.default_enabled,
            params=self.default_parameters(),
            execution=self.pipeline_execution_metadata(),
            supports_tiled_input=self.supports_tiled_input(),
            stage=m.stage,
        )


def binding():
    """(ModuleBase, ModuleMetadata, ModuleStage, StepExecutionMetadata, PipelineStep) to subclass.

    Inside the reference application (its packages are importable) its own classes are returned;
    stand-alone, this package's mirror.
    """
    try:  # pragma: no cover - exercised only with the reference checkout on sys.path
        from plugins.module_base import ModuleBase as RB, ModuleMetadata as RM, ModuleStage as RS
        from processing.pipeline_manager import PipelineStep as RP, StepExecutionMetadata as RE

        return RB, RM, RS, RE, RP
    except Exception:
        return ModuleBase, ModuleMetadata, ModuleStage, StepExecutionMetadata, PipelineStep


__all__ = ["MenuEntry", "ModuleBase", "ModuleMetadata", "ModuleStage", "binding"]
