#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference checkout and cv2 4.13.0 live there):

    python tests/golden/make_golden.py

Every array in the fixtures is the output of a reference function (or, for the ops the reference
lacks, of the same third-party call the reference would make) on a small seeded input that is
stored next to it.  PyQt5 / skimage / skfuzzy / sklearn are stubbed exactly as SURVEY.md App. C
describes; no reference source is copied.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))


from oracle.ref_stubs import install_stubs  # noqa: E402


def main() -> None:
    install_stubs()
    sys.path.insert(0, str(REF))
    import cv2
    from modules import preprocessing as mp
    from core import segmentation as cs
    from core import preprocessing as cp
    from processing.pipeline_manager import PipelineManager, PipelineStep
    from processing.pipeline_cache import PipelineCache

    rng = np.random.default_rng(20261018)
    H, W = 48, 64

    def rnd(shape, dt):
        return rng.integers(0, (255 if dt == np.uint8 else 65535) + 1, shape, dtype=dt)

    def smooth(dt):
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.zeros((H, W))
        for _ in range(9):
            cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(3, 8)
            img += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * r * r)) * rng.uniform(0.4, 1.0)
        img = np.clip(img, 0, 1) * 0.6 + 0.05 + rng.normal(0, 0.01, (H, W))
        hi = 255 if dt == np.uint8 else 65535
        return np.clip(img * hi, 0, hi).astype(dt)

    g: dict[str, np.ndarray] = {}
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "reference": str(REF), "seed": 20261018}

    # ---- preprocessing modules (modules/preprocessing.py) ----
    for tag, dt in (("u8", np.uint8), ("u16", np.uint16)):
        bgr = rnd((H, W, 3), dt)
        gray = smooth(dt)
        noisy = rnd((H, W), dt)
        g[f"in_bgr_{tag}"], g[f"in_gray_{tag}"], g[f"in_noise_{tag}"] = bgr, gray, noisy
        g[f"grayscale_{tag}"] = mp.GrayscaleModule().process(bgr)
        for k in (3, 5, 11, 15):
            g[f"gauss{k}_{tag}"] = mp.NoiseReductionModule().process(noisy, method="Gaussian", ksize=k)
        for k in (3, 5):
            g[f"median{k}_{tag}"] = mp.NoiseReductionModule().process(noisy, method="Median", ksize=k)
        g[f"normalize_{tag}"] = mp.IntensityNormalizationModule().process(np.maximum(gray, dt(9)), alpha=0, beta=255)
        g[f"normalize_10_200_{tag}"] = mp.IntensityNormalizationModule().process(np.maximum(gray, dt(9)), alpha=10, beta=200)
        g[f"brightness_{tag}"] = mp.BrightnessContrastModule().process(noisy, alpha=1.5, beta=-20)
        # ---- segmentation functions (core/segmentation.py) ----
        g[f"otsu_{tag}"] = cs.otsu_threshold(gray)
        g[f"otsu_bgr_{tag}"] = cs.otsu_threshold(bgr)
        g[f"global100_{tag}"] = cs.global_threshold(gray, 100)
        for shape in ("Rectangular", "Elliptical", "Cross"):
            for k, it in ((3, 1), (5, 2)):
                key = f"{shape.lower()}_{k}_{it}_{tag}"
                g[f"open_{key}"] = cs.morphological_opening(noisy, shape, k, it)
                g[f"close_{key}"] = cs.morphological_closing(noisy, shape, k, it)
                g[f"dilate_{key}"] = cs.morphological_dilation(noisy, shape, k, it)
                g[f"erode_{key}"] = cs.morphological_erosion(noisy, shape, k, it)
        # ops the reference lacks: same library, called the way the reference would
        g[f"clahe_{tag}"] = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(gray)
        g[f"clahe_4_3x5_{tag}"] = cv2.createCLAHE(clipLimit=4.0, tileGridSize=(3, 5)).apply(gray)
        g[f"box5_{tag}"] = cv2.blur(noisy, (5, 5))
    g["gamma22_u8"] = mp.GammaCorrectionModule().process(g["in_noise_u8"], gamma=2.2)
    g["adaptive_11_2_u8"] = cs.Detector.adaptive_threshold(g["in_gray_u8"], 11, 2)
    g["adaptive_31_m3_u8"] = cs.Detector.adaptive_threshold(g["in_gray_u8"], 31, -3)
    g["adaptive_bgr_u8"] = cs.Detector.adaptive_threshold(g["in_bgr_u8"], 11, 2)
    g["equalize_u8"] = cp.Preprocessor.histogram_equalization(g["in_gray_u8"])
    # connected components: the reference's cv2 call (core/segmentation.py:108) + stats
    mask = cs.otsu_threshold(g["in_gray_u8"])
    n, lab, stats, cent = cv2.connectedComponentsWithStats(mask, connectivity=8)
    g["ccl_mask_u8"], g["ccl_labels_cv2"], g["ccl_stats_cv2"], g["ccl_centroids_cv2"] = mask, lab, stats, cent
    meta["ccl_count"] = int(n - 1)

    # ---- driver semantics (processing/pipeline_manager.py) ----
    def add(image, *, value): return image + value
    def mul(image, *, factor): return image * factor
    base = np.arange(16, dtype=np.float32).reshape(4, 4)
    pm = PipelineManager([PipelineStep("add", add, params={"value": 1.5}), PipelineStep("mul", mul, params={"factor": 2.0})])
    g["pm_in"], g["pm_out"] = base, pm.apply(base)
    stack = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    g["pm_stack_in"], g["pm_stack_out"] = stack, pm.apply(stack)

    # ---- cache-key contract (processing/pipeline_cache.py:256-313) ----
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        cache = PipelineCache()
        if hasattr(cache, "set_cache_directory"):
            cache.set_cache_directory(Path(td))
        a = np.arange(16, dtype=np.uint16).reshape(4, 4)
        sid = cache.register_source(a)
        steps = [
            PipelineStep("Grayscale", lambda x: x, enabled=True, params={}),
            PipelineStep("NoiseReduction", lambda x, **k: x, enabled=True, params={"method": "Gaussian", "ksize": 5}),
            PipelineStep("IntensityNormalization", lambda x, **k: x, enabled=False, params={"alpha": 0, "beta": 255}),
        ]
        final, records = cache.predict(sid, steps)
        meta["cache"] = {"source_id": sid, "final": final, "signatures": [r.signature for r in records]}
        steps2 = [PipelineStep("Opening", lambda x, **k: x, params={"kernel_shape": "Rectangular", "kernel_size": 5, "iterations": 1}),
                  PipelineStep("X", lambda x, **k: x, params={"t": (1, 2), "m": {"b": 1, "a": [2.5, None]}})]
        final2, records2 = cache.predict(sid, steps2)
        meta["cache2"] = {"final": final2, "signatures": [r.signature for r in records2]}

    np.savez_compressed(OUT / "reference_outputs.npz", **g)
    (OUT / "reference_meta.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
    print(f"wrote {len(g)} arrays, {(OUT / 'reference_outputs.npz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
