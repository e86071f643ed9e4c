"""CPU reference path: the reference's call sites restated on the SAME third-party library.

TEST / BENCHMARK INFRASTRUCTURE ONLY (``bench.py --impl reference`` and ``cpu_baseline``; never
imported by ``yamimageprocessor_b200``).  The reference cannot travel to the GPU box
(``/root/reference`` is absent there), but what it executes for this path is a handful of
``cv2`` / ``scipy`` calls; each function below makes exactly the call the cited reference line
makes, with ``cv2.setNumThreads(os.cpu_count())``.  When cv2 is not importable the NumPy oracle
(``np_oracle``) is used instead and ``KIND`` says so.

Ops the reference lacks or cannot run on uint16 follow SURVEY.md §8(c): CLAHE -> cv2.createCLAHE;
uint16 adaptive threshold -> cv2.GaussianBlur(float32, BORDER_REPLICATE) + rint + compare (what
cv2.adaptiveThreshold does internally for uint8); labels -> scipy.ndimage.label with the 3x3
structure (identical numbering to skimage.measure.label); region table -> np.bincount.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import np_oracle as O

try:  # pragma: no cover - depends on the image
    import cv2  # type: ignore

    cv2.setNumThreads(os.cpu_count() or 1)
    HAVE_CV2 = True
    KIND = f"cv2 {cv2.__version__} + scipy (reference call sites restated)"
    THREADS = cv2.getNumThreads()
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False
    KIND = "numpy oracle (cv2 not importable)"
    THREADS = 1

_SHAPES = {"rectangular": 0, "elliptical": 2, "cross": 1}  # cv2.MORPH_RECT / ELLIPSE / CROSS


def to_grayscale(img):  # modules/preprocessing.py:52-55
    if img.ndim == 3:
        return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if HAVE_CV2 else O.bgr2gray(img)
    return img


def noise_reduction_gaussian(img, ksize):  # modules/preprocessing.py:145
    return cv2.GaussianBlur(img, (ksize, ksize), 0) if HAVE_CV2 else O.gaussian(img, ksize, 0.0)


def clahe(img, clip_limit=2.0, grid=(8, 8)):  # absent from the reference; same library
    if HAVE_CV2:
        return cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=grid).apply(img)
    return O.clahe(img, clip_limit, grid)


def otsu_threshold(img):  # core/segmentation.py:145-148
    gray = to_grayscale(img)
    if HAVE_CV2 and gray.size < (1 << 31):
        return cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
    return O.otsu_threshold(gray, 255)[1]


def adaptive_threshold(img, block_size=11, C=2):  # core/segmentation.py:91-94
    gray = to_grayscale(img)
    if not HAVE_CV2:
        return O.adaptive_threshold(gray, block_size, C)
    if gray.dtype == np.uint8:
        return cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, block_size, C)
    # uint16 extension: what adaptiveThreshold does internally, on the 16-bit range
    mean = cv2.GaussianBlur(gray.astype(np.float32), (block_size, block_size), 0, borderType=cv2.BORDER_REPLICATE)
    mean = np.clip(np.rint(mean), 0, 65535).astype(np.int32)
    return np.where(gray.astype(np.int32) - mean > -int(math.ceil(C)), 255, 0).astype(np.uint8)


def _kernel(kernel_shape, kernel_size):  # core/segmentation.py:265-274
    return cv2.getStructuringElement(_SHAPES.get(kernel_shape.lower(), 0), (kernel_size, kernel_size))


def morphological_opening(img, kernel_shape="Rectangular", kernel_size=3, iterations=1):  # core/segmentation.py:264-275
    if not HAVE_CV2:
        return O.morph_open(img, kernel_shape, kernel_size, iterations)
    return cv2.morphologyEx(img, cv2.MORPH_OPEN, _kernel(kernel_shape, kernel_size), iterations=iterations)


def morphological_closing(img, kernel_shape="Rectangular", kernel_size=3, iterations=1):  # core/segmentation.py:277-288
    if not HAVE_CV2:
        return O.morph_close(img, kernel_shape, kernel_size, iterations)
    return cv2.morphologyEx(img, cv2.MORPH_CLOSE, _kernel(kernel_shape, kernel_size), iterations=iterations)


def label(mask):  # core/extraction.py:60,73 (skimage.measure.label == scipy.ndimage.label, 8-conn)
    return O.ccl_label(mask)[1]


def connected_components(mask):  # core/segmentation.py:108 -- the SEGMENTATION call site (8-conn, int32)
    if not HAVE_CV2:
        return O.ccl_label(mask)[1]
    return cv2.connectedComponents(mask)[1]


def region_table(labels, intensity):  # core/extraction.py:74-87 (regionprops restated, + mean intensity)
    return O.region_props(labels, intensity)


# ---- the BASELINE.json configurations as whole pipelines --------------------------------------
def preprocess(frame):
    """C1: gray -> Gaussian sigma=2 (NoiseReduction ksize=11) -> CLAHE(2.0, 8x8) -> Otsu."""
    g = to_grayscale(frame)
    g = noise_reduction_gaussian(g, 11)
    g = clahe(g, 2.0, (8, 8))
    return g, otsu_threshold(g)


def segment(frame):
    """C2: adaptive threshold (11, 2) -> open 5x5 -> close 5x5 -> 8-connected labels
    (cv2.connectedComponents, the segmentation call site; same partition as the raster-first numbering)."""
    m = adaptive_threshold(frame, 11, 2)
    m = morphological_opening(m, "Rectangular", 5, 1)
    m = morphological_closing(m, "Rectangular", 5, 1)
    return connected_components(m)


def extract(labels, intensity):
    """C3: per-region area / centroid / bbox / mean intensity."""
    return region_table(labels, intensity)


def full_chain(frame):
    """C4/C5 chain: preprocess, then segment the CLAHE output, then extract."""
    g, otsu_mask = preprocess(frame)
    lab = segment(g)
    return otsu_mask, lab, extract(label((lab > 0).astype(np.uint8)), g)   # extraction labels: core/extraction.py:73


def mosaic_chain(frame):
    """C4 chain: preprocess, then segment the CLAHE output (no extraction)."""
    g, otsu_mask = preprocess(frame)
    return otsu_mask, segment(g)
