"""Parameter registry of the GPU-backed modules.

Defaults, ranges and coercions restate the hot-path entries of the reference's
``ui/control_metadata.py`` (``ControlMetadata.coerce:93-125``, ``_ensure_odd:132``,
``MODULE_CONTROL_METADATA:146-687``) so ``sanitize_parameters`` behaves the same without importing
the UI package.  Parameter NAMES are the reference's: they feed the ``pipeline_cache``
signatures (``processing/pipeline_cache.py:291-313``), so they must not change.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple


def ensure_odd(value: Any) -> int:
    n = int(round(float(value)))
    return n + 1 if n % 2 == 0 else n


@dataclass(frozen=True)
class ParamSpec:
    default: Any = None
    kind: str = "float"  # "int" | "float" | "bool" | "str"
    minimum: Optional[float] = None
    maximum: Optional[float] = None
    decimals: Optional[int] = None
    choices: Tuple[Any, ...] = ()
    coerce_fn: Optional[Callable[[Any], Any]] = None

    def coerce(self, value: Any) -> Any:
        if value is None:
            return self.default
        try:
            if self.kind == "int":
                value = int(round(float(value)))
            elif self.kind == "float":
                value = float(value)
            elif self.kind == "bool":
                value = bool(value)
            elif self.kind == "str":
                value = str(value)
        except (TypeError, ValueError):
            return self.default
        if self.coerce_fn is not None:
            value = self.coerce_fn(value)
        if isinstance(value, (int, float)) and not isinstance(value, bool):
            if self.minimum is not None:
                value = max(value, self.minimum)
            if self.maximum is not None:
                value = min(value, self.maximum)
            if self.kind == "int":
                value = int(round(value))
            elif self.kind == "float" and self.decimals is not None:
                value = round(float(value), self.decimals)
        if self.choices and value not in self.choices:
            return self.default if self.default is not None else self.choices[0]
        return value


_SHAPES = ("Rectangular", "Elliptical", "Cross")


def _morph() -> Dict[str, ParamSpec]:
    return {
        "kernel_shape": ParamSpec("Rectangular", "str", choices=_SHAPES),
        "kernel_size": ParamSpec(3, "int", 1, 31),
        "iterations": ParamSpec(1, "int", 1, 10),
    }


MODULE_PARAMS: Dict[str, Dict[str, ParamSpec]] = {
    # ---- preprocessing (modules/preprocessing.py) ----
    "Grayscale": {},
    "BrightnessContrast": {
        "alpha": ParamSpec(1.0, "float", 0.1, 3.0, decimals=2),
        "beta": ParamSpec(0, "int", -100, 100),
    },
    "Gamma": {"gamma": ParamSpec(1.0, "float", 0.1, 5.0, decimals=2)},
    "IntensityNormalization": {
        "alpha": ParamSpec(0, "int", 0, 255),
        "beta": ParamSpec(255, "int", 0, 255),
    },
    "NoiseReduction": {
        "method": ParamSpec("Gaussian", "str", choices=("Gaussian", "Median", "Bilateral")),
        "ksize": ParamSpec(5, "int", 1, 15, coerce_fn=ensure_odd),
    },
    "Sharpen": {"strength": ParamSpec(1.0, "float", 0.0, 5.0, decimals=2)},
    "SelectChannel": {"channel": ParamSpec("All", "str", choices=("All", "R", "G", "B", "RG", "GB", "BR"))},
    # ---- north_star ops the reference lacks (same third-party library, cv2 semantics) ----
    "CLAHE": {
        "clip_limit": ParamSpec(2.0, "float", 0.0, 40.0, decimals=2),
        "tile_grid_x": ParamSpec(8, "int", 1, 64),
        "tile_grid_y": ParamSpec(8, "int", 1, 64),
    },
    "BoxFilter": {"ksize": ParamSpec(3, "int", 1, 31, coerce_fn=ensure_odd)},
    "HistogramEqualization": {},
    # ---- segmentation (core/segmentation.py via processing/segmentation_pipeline.py:84-184) ----
    "Global": {"threshold": ParamSpec(127, "int", 0, 255)},
    "Otsu": {},
    "Adaptive": {
        "block_size": ParamSpec(11, "int", 3, 101, coerce_fn=ensure_odd),
        "C": ParamSpec(2, "int", -10, 10),
    },
    "Sobel": {"ksize": ParamSpec(3, "int", 1, 31, coerce_fn=ensure_odd)},
    "Prewitt": {},
    "Laplacian": {"ksize": ParamSpec(3, "int", 1, 31, coerce_fn=ensure_odd)},
    "Border Removal": {"border_distance": ParamSpec(100, "int", 0, 999)},
    "Opening": _morph(),
    "Closing": _morph(),
    "Dilation": _morph(),
    "Erosion": _morph(),
    "ConnectedComponents": {},
    # ---- extraction (core/extraction.py:57-87) ----
    # (labels of the Otsu mask; the reference's "Region Properties" STEP returns an annotated copy of its
    #  input, core/extraction.py:57-68, so the label-producing step carries its own name and signature)
    "RegionLabels": {},
    # ---- BASELINE config 4: the whole chain over a tiled handle (supports_tiled_input) ----
    "Mosaic": {
        "gauss_ksize": ParamSpec(11, "int", 1, 15, coerce_fn=ensure_odd),
        "clip_limit": ParamSpec(2.0, "float", 0.0, 40.0, decimals=2),
        "tile_grid_x": ParamSpec(8, "int", 1, 64),
        "tile_grid_y": ParamSpec(8, "int", 1, 64),
        "block_size": ParamSpec(11, "int", 3, 15, coerce_fn=ensure_odd),
        "C": ParamSpec(2, "int", -10, 10),
        "morph_ksize": ParamSpec(5, "int", 1, 31, coerce_fn=ensure_odd),
        "strips": ParamSpec(1, "int", 1, 64),
    },
}

__all__ = ["MODULE_PARAMS", "ParamSpec", "ensure_odd"]
