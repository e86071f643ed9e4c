"""Reference arm of the benchmark: the UNMODIFIED reference when its checkout is importable, else the
restated call sites of ``cv2_path``.

TEST / BENCHMARK INFRASTRUCTURE ONLY (``bench.py --impl reference`` / ``cpu_baseline``); never imported
by ``yamimageprocessor_b200``.  ``/root/reference`` exists in the build container only; on the GPU box
the port runs (``KIND`` says which).  With the reference present the pipelines are built from its own
classes and functions and driven by its own ``PipelineManager.apply``:

* ``modules.preprocessing.NoiseReductionModule.process`` (``modules/preprocessing.py:140-150``),
  ``core.segmentation.otsu_threshold`` (``:145-148``), ``morphological_opening`` / ``_closing``
  (``:264-288``) -- imported, not restated;
* the steps the reference lacks or cannot run on uint16 (CLAHE, uint16 adaptive
  threshold, a stand-alone connected-components step) use the same third-party calls as ``cv2_path``
  (``cv2.createCLAHE``, float32 ``cv2.GaussianBlur`` + compare, ``cv2.connectedComponents`` -- the call
  ``core/segmentation.py:108`` makes).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

from . import cv2_path as P

REF = Path(os.environ.get("YAM_REFERENCE", "/root/reference"))
KIND = "port"
DETAIL = P.KIND
THREADS = P.THREADS
_pipelines = None


def _load():
    """(preprocess_pm, segment_pm) built from the reference's own code, or None."""
    global _pipelines, KIND, DETAIL
    if _pipelines is not None:
        return _pipelines or None
    _pipelines = False
    if not (REF / "processing" / "pipeline_manager.py").exists() or not P.HAVE_CV2:
        return None
    try:
        from .ref_stubs import install_stubs

        install_stubs()
        if str(REF) not in sys.path:
            sys.path.insert(0, str(REF))
        from core import segmentation as cs
        from modules import preprocessing as mp
        from processing.pipeline_manager import PipelineManager, PipelineStep

        def step(name, fn, **params):
            return PipelineStep(name=name, function=fn, enabled=True, params=params)

        pre = PipelineManager([
            step("NoiseReduction", mp.NoiseReductionModule().process, method="Gaussian", ksize=11),
            step("CLAHE", lambda img: P.clahe(img, 2.0, (8, 8))),
        ])
        otsu = PipelineManager([step("Otsu", cs.otsu_threshold)])
        seg = PipelineManager([
            step("Adaptive", lambda img: P.adaptive_threshold(img, 11, 2)),
            step("Opening", cs.morphological_opening, kernel_shape="Rectangular", kernel_size=5, iterations=1),
            step("Closing", cs.morphological_closing, kernel_shape="Rectangular", kernel_size=5, iterations=1),
            step("ConnectedComponents", P.connected_components),
        ])
        _pipelines = (pre, otsu, seg)
        KIND = "reference"
        DETAIL = f"unmodified reference from {REF} (PipelineManager.apply + its own step functions; {P.KIND} for the ops it lacks)"
    except Exception as exc:  # pragma: no cover - depends on the checkout
        DETAIL = f"{P.KIND}; importing the reference failed: {exc!r}"
        _pipelines = False
        return None
    return _pipelines


def preprocess(frame):
    pm = _load()
    if not pm:
        return P.preprocess(frame)
    g = pm[0].apply(frame)
    return g, pm[1].apply(g)


def segment(frame):
    pm = _load()
    return pm[2].apply(frame) if pm else P.segment(frame)


def extract(labels, intensity):
    return P.extract(labels, intensity)


def mosaic_chain(frame):
    g, otsu_mask = preprocess(frame)
    return otsu_mask, segment(g)


def full_chain(frame):
    import numpy as np

    g, otsu_mask = preprocess(frame)
    lab = segment(g)
    return otsu_mask, lab, extract(P.label((lab > 0).astype(np.uint8)), g)


def describe():
    _load()
    return KIND, DETAIL
