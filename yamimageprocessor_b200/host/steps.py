"""Step name -> device operator table.

Keys are the step names the reference uses (they enter the ``pipeline_cache`` signatures):
preprocessing module identifiers (``modules/preprocessing.py:46,66,89,113,134,161,182``), segmentation
method names (``processing/segmentation_pipeline.py:84-184``), plus the north_star ops the reference
lacks (``CLAHE``, ``BoxFilter``, ``HistogramEqualization``, ``ConnectedComponents``, ``RegionLabels``).
The reference's extraction STEP ``Region Properties`` (``processing/extraction_pipeline.py:84-86``)
returns an annotated copy of its input (``cv2.rectangle`` / ``cv2.circle`` drawing, UI); it is not a
device step -- a step of that name returning labels would share its cache signature while holding
different content.  The numbers behind it come from ``region_properties_data``.

Every function takes ``(backend, tensor, params)`` with a CUDA tensor shaped ``(h, w)``,
``(n, h, w)`` (stack, processed per frame like ``_apply_slice_wise``) or ``(h, w, 3)`` BGR and
returns a CUDA tensor.  Parameter names and defaults follow the reference call sites.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Mapping

import numpy as np

from ..backend import Backend
from .._lib import MORPH_CLOSE, MORPH_DILATE, MORPH_ERODE, MORPH_OPEN


class UnsupportedOnDevice(NotImplementedError):
    """Raised for inputs/parameters the hot path does not cover (there is no CPU fallback)."""


def _is_colour(t) -> bool:
    return t.dim() == 3 and t.shape[-1] in (3, 4)


def _gray(be: Backend, t):
    """Preprocessor.to_grayscale (core/segmentation.py:47-48): BGR2GRAY when 3-D colour."""
    return be.bgr2gray(t) if _is_colour(t) else t


def _plane_only(t, name: str):
    if _is_colour(t):
        raise UnsupportedOnDevice(
            f"{name}: colour input is outside the GPU hot path; run Grayscale first (single-channel planes only)"
        )
    return t


def _per_channel(be: Backend, t, fn):
    """cv2's neighbourhood filters treat the channels of an interleaved colour image independently
    (modules/preprocessing.py:140-150 is channel-agnostic): run ``fn`` on the (c, h, w) plane stack."""
    if not _is_colour(t):
        return fn(t)
    return be.merge_channels(fn(be.split_channels(t)))


def grayscale(be: Backend, t, p: Mapping[str, Any]):
    return _gray(be, t)


def brightness_contrast(be: Backend, t, p):
    alpha = float(p.get("alpha", 1.0))
    if alpha <= 0:
        raise ValueError("Alpha must be > 0")  # modules/preprocessing.py:76
    return be.convert_scale_abs(t, alpha, float(p.get("beta", 0)))


def gamma(be: Backend, t, p):
    g = float(p.get("gamma", 1.0))
    if g <= 0:
        raise ValueError("Gamma must be > 0")  # modules/preprocessing.py:98
    inv = 1.0 / g
    table = np.array([(i / 255.0) ** inv * 255 for i in range(256)]).astype("uint8")
    return be.lut_u8(t, table)  # uint8 only, like cv2.LUT


def intensity_normalization(be: Backend, t, p):
    if _is_colour(t):
        # cv2.normalize(NORM_MINMAX) on a multi-channel image takes ONE min / max over all channels
        # (cv::norm / minMaxLoc on the reshaped matrix): the interleaved frame is one (h, w*c) plane
        h, w, c = (int(v) for v in t.shape)
        return be.normalize_minmax(t.reshape(h, w * c), float(p.get("alpha", 0)), float(p.get("beta", 255))).reshape(h, w, c)
    return be.normalize_minmax(t, float(p.get("alpha", 0)), float(p.get("beta", 255)))


def noise_reduction(be: Backend, t, p):
    method = str(p.get("method", "Gaussian"))
    ksize = int(p.get("ksize", 5))
    if method == "Gaussian":
        return _per_channel(be, t, lambda x: be.gaussian(x, ksize, 0.0))
    if method == "Median":
        return _per_channel(be, t, lambda x: be.median(x, ksize))
    if method == "Bilateral":
        raise UnsupportedOnDevice("NoiseReduction(method='Bilateral') is outside the GPU hot path")
    return t  # unknown method: identity, like modules/preprocessing.py:150


def clahe(be: Backend, t, p):
    grid = (int(p.get("tile_grid_x", 8)), int(p.get("tile_grid_y", 8)))
    return be.clahe(_plane_only(t, "CLAHE"), float(p.get("clip_limit", 2.0)), grid)


def box_filter(be: Backend, t, p):
    return _per_channel(be, t, lambda x: be.box(x, int(p.get("ksize", 3))))


def histogram_equalization(be: Backend, t, p):
    if _is_colour(t):   # core/preprocessing.py:77-79: equalise Y of YCrCb, convert back
        return be.equalize_hist_bgr(t)
    return be.equalize_hist(t)


def global_threshold(be: Backend, t, p):
    return be.threshold(_gray(be, t), float(p.get("threshold", 127)), 255)


def otsu(be: Backend, t, p):
    return be.otsu_threshold(_gray(be, t), 255)[1]


def adaptive(be: Backend, t, p):
    return be.adaptive_threshold(_gray(be, t), int(p.get("block_size", 11)), float(p.get("C", 2)))


def sharpen(be: Backend, t, p):
    return _per_channel(be, t, lambda x: be.sharpen(x, float(p.get("strength", 1.0))))


def select_channel(be: Backend, t, p):
    if t.dim() == 3 and t.shape[-1] != 3:
        raise UnsupportedOnDevice("SelectChannel: stacks / 4-channel images are outside the GPU hot path")
    return be.select_channel(t, str(p.get("channel", "All")))


def _edge(kind: str) -> Callable:
    def run(be: Backend, t, p):
        ksize = int(p.get("ksize", 3))
        if ksize not in (1, 3, 5, 7):
            raise UnsupportedOnDevice(f"{kind}: ksize {ksize} exceeds the exact integer range of the device kernel (1, 3, 5, 7)")
        return be.edge_filter(_gray(be, t), kind, ksize)

    return run


def border_removal(be: Backend, t, p):
    return be.border_clear(t, int(p.get("border_distance", 100)))


def _morph(op: int) -> Callable:
    def run(be: Backend, t, p):
        return _per_channel(be, t, lambda x: be.morph(
            x,
            op,
            str(p.get("kernel_shape", "Rectangular")),
            int(p.get("kernel_size", 3)),
            int(p.get("iterations", 1)),
        ))

    return run


def _as_mask_u8(be: Backend, t):
    import torch

    if t.dtype == torch.uint8:
        return t
    if t.dtype == torch.uint16:
        return be.convert_scale_abs(t, 1.0, 0.0)  # saturates: non-zero stays non-zero
    raise UnsupportedOnDevice(f"ConnectedComponents: mask dtype {t.dtype} not supported")


def connected_components(be: Backend, t, p):
    return be.ccl_label(_as_mask_u8(be, _plane_only(t, "ConnectedComponents")))[0]


def region_properties_labels(be: Backend, t, p):
    """Otsu -> 8-connected labels (core/extraction.py:58-60); the table comes from region_table()."""
    mask = be.otsu_threshold(_gray(be, t), 255)[1]
    return be.ccl_label(_as_mask_u8(be, mask))[0]


def mosaic_dense(be: Backend, t, p):
    """The ``Mosaic`` chain on a frame that is already resident (one strip, no collectives): Gaussian ->
    CLAHE -> adaptive threshold -> open -> close -> labels.  The lazy-handle entry is MosaicModule.process."""
    from . import mosaic

    mp = mosaic.MosaicParams(gauss_ksize=int(p.get("gauss_ksize", 11)), clip_limit=float(p.get("clip_limit", 2.0)),
                             tile_grid=(int(p.get("tile_grid_x", 8)), int(p.get("tile_grid_y", 8))),
                             block_size=int(p.get("block_size", 11)), C=float(p.get("C", 2)),
                             morph_ksize=int(p.get("morph_ksize", 5)))
    t = _plane_only(t, "Mosaic")
    if t.dim() != 2:
        raise UnsupportedOnDevice("Mosaic: one (H, W) frame at a time")
    return mosaic.run_strip(be, t, 0, 1, mp, device_source=t).labels


DEVICE_STEPS: Dict[str, Callable] = {
    "Grayscale": grayscale,
    "BrightnessContrast": brightness_contrast,
    "Gamma": gamma,
    "IntensityNormalization": intensity_normalization,
    "NoiseReduction": noise_reduction,
    "CLAHE": clahe,
    "BoxFilter": box_filter,
    "HistogramEqualization": histogram_equalization,
    "Sharpen": sharpen,
    "SelectChannel": select_channel,
    "Sobel": _edge("sobel"),
    "Prewitt": _edge("prewitt"),
    "Laplacian": _edge("laplacian"),
    "Border Removal": border_removal,
    "Global": global_threshold,
    "Otsu": otsu,
    "Adaptive": adaptive,
    "Opening": _morph(MORPH_OPEN),
    "Closing": _morph(MORPH_CLOSE),
    "Dilation": _morph(MORPH_DILATE),
    "Erosion": _morph(MORPH_ERODE),
    "ConnectedComponents": connected_components,
    "RegionLabels": region_properties_labels,
    "Mosaic": mosaic_dense,
}


def region_table(be: Backend, labels, intensity=None, n_labels=None) -> Dict[str, np.ndarray]:
    """Per-region table for a single labelled frame, columns in skimage semantics
    (core/extraction.py:70-87: region_index = label, centroid = (row, col)); bbox is half-open."""
    from ..backend import contour_columns, props_table, shape_columns

    if n_labels is None:
        n_labels = int(labels.max().item()) if labels.numel() else 0
    props_dev = be.region_props(labels, intensity, n_labels)
    props = be.to_host(props_dev)
    table = props_table(props)
    # extent / eccentricity / orientation (core/extraction.py:81-85) from exact second-order sums
    table.update(shape_columns(props, be.to_host(be.region_moments(labels, n_labels))))
    # perimeter / solidity (core/extraction.py:80,83) from exact border-class counts and hull pixel counts
    table.update(contour_columns(props, be.to_host(be.region_perimeter_counts(labels, n_labels)),
                                 be.to_host(be.region_convex_area(labels, n_labels, props_dev))))
    table["region_index"] = np.arange(1, n_labels + 1, dtype=np.int64)
    table["centroid"] = np.stack([table["centroid_row"], table["centroid_col"]], axis=1) if n_labels else np.zeros((0, 2))
    if intensity is None:
        table.pop("mean_intensity", None)
        table.pop("sum_intensity", None)
    return table


__all__ = ["DEVICE_STEPS", "UnsupportedOnDevice", "region_table"]
