#!/usr/bin/env python
"""Golden vectors for SURVEY.md 8(f) N4 (watershed front half), produced by the UNMODIFIED reference
from /root/reference plus the cv2 calls its function makes internally.

    python tests/golden/make_golden_n4.py      # build container only (needs the reference + cv2 4.13.0)

``Detector.watershed_segmentation`` (core/segmentation.py:97-114) returns the input with the watershed
lines painted red; the marker construction it performs on the way (:99-110) is reproduced here line by
line with the same cv2 calls so that every intermediate is pinned: thresh, opening, sure_bg, dist
(cv2.distanceTransform DIST_L2 / 5), sure_fg and the markers image handed to cv2.watershed.  The
reference's own output is stored too; it equals cv2.watershed applied to those markers (asserted).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as base  # noqa: E402

OUT = Path(__file__).resolve().parent


def cells(rng, h, w, n):
    """dark blobs on a bright noisy background (so that THRESH_BINARY_INV + OTSU selects the blobs)"""
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.full((h, w), 200.0)
    for _ in range(n):
        cy, cx, r = rng.integers(6, h - 6), rng.integers(6, w - 6), rng.integers(3, 9)
        img -= 150.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * r * r))
    return np.clip(img + rng.normal(0, 4, (h, w)), 0, 255).astype(np.uint8)


def main() -> None:
    base.install_stubs()
    sys.path.insert(0, str(base.REF))
    import cv2
    from core import segmentation as cs

    rng = np.random.default_rng(20261020)
    g: dict[str, np.ndarray] = {}
    for i, (h, w, n, params) in enumerate(((96, 128, 14, {}), (70, 90, 9, dict(kernel_size=5, opening_iterations=1, dilation_iterations=2,
                                                                                  distance_threshold_factor=0.5)),
                                           (120, 100, 20, dict(kernel_size=3, opening_iterations=1, dilation_iterations=1,
                                                               distance_threshold_factor=0.35)))):
        gray = cells(rng, h, w, n)
        bgr = np.ascontiguousarray(np.stack([gray, np.clip(gray.astype(int) + 3, 0, 255).astype(np.uint8), gray], axis=-1))
        p = dict(kernel_size=3, opening_iterations=2, dilation_iterations=3, distance_threshold_factor=0.7)
        p.update(params)
        g[f"in_bgr_{i}"] = bgr
        g[f"params_{i}"] = np.array([p["kernel_size"], p["opening_iterations"], p["dilation_iterations"], p["distance_threshold_factor"]])
        g[f"watershed_{i}"] = cs.Detector.watershed_segmentation(bgr, **p)       # the reference step's output
        # the same lines, intermediates kept (core/segmentation.py:99-110)
        gr = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        ret, thresh = cv2.threshold(gr, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)
        kernel = np.ones((p["kernel_size"], p["kernel_size"]), np.uint8)
        opening = cv2.morphologyEx(thresh, cv2.MORPH_OPEN, kernel, iterations=p["opening_iterations"])
        sure_bg = cv2.dilate(opening, kernel, iterations=p["dilation_iterations"])
        dist = cv2.distanceTransform(opening, cv2.DIST_L2, 5)
        ret, sure_fg = cv2.threshold(dist, p["distance_threshold_factor"] * dist.max(), 255, 0)
        sure_fg = np.uint8(sure_fg)
        unknown = cv2.subtract(sure_bg, sure_fg)
        ret, markers = cv2.connectedComponents(sure_fg)
        markers = markers + 1
        markers[unknown == 255] = 0
        flooded = cv2.watershed(bgr, markers.copy())
        annotated = bgr.copy()
        annotated[flooded == -1] = [0, 0, 255]
        assert np.array_equal(annotated, g[f"watershed_{i}"])   # the front half reproduced here IS what the reference feeds cv2.watershed
        g[f"thresh_{i}"] = thresh
        for name, arr in (("gray", gr), ("opening", opening), ("sure_bg", sure_bg), ("dist", dist), ("sure_fg", sure_fg), ("markers", markers)):
            g[f"{name}_{i}"] = arr
    np.savez_compressed(OUT / "reference_outputs_n4.npz", **g)
    print(f"wrote {len(g)} arrays, {(OUT / 'reference_outputs_n4.npz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
