#!/usr/bin/env python
"""Golden vectors for histogram equalisation of COLOUR images (core/preprocessing.py:74-79: equalise the
Y plane of YCrCb), produced by the UNMODIFIED reference from /root/reference.

    python tests/golden/make_golden_color.py      # build container only (needs the reference + cv2 4.13.0)

Same stubbing as make_golden.py; inputs are stored next to the outputs in reference_outputs_color.npz.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as base  # noqa: E402

OUT = Path(__file__).resolve().parent


def main() -> None:
    base.install_stubs()
    sys.path.insert(0, str(base.REF))
    from core.preprocessing import Preprocessor

    rng = np.random.default_rng(20261020)
    g: dict[str, np.ndarray] = {}
    yy, xx = np.mgrid[0:61, 0:83]
    smooth = np.stack([np.clip((np.sin(yy / (4.0 + c)) * np.cos(xx / (6.0 + 2 * c)) * 0.35 + 0.45 + 0.1 * c) * 255
                               + rng.normal(0, 6, yy.shape), 0, 255) for c in range(3)], axis=-1).astype(np.uint8)
    cases = {
        "noise": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
        "smooth": smooth,
        "dark": (rng.integers(0, 40, (24, 31, 3))).astype(np.uint8),
        "saturated": np.where(rng.random((29, 30, 3)) < 0.5, 255, rng.integers(0, 256, (29, 30, 3))).astype(np.uint8),
        "constant": np.full((9, 11, 3), (10, 200, 77), np.uint8),
    }
    for name, img in cases.items():
        g[f"in_{name}"] = img
        g[f"equalized_{name}"] = Preprocessor.histogram_equalization(img.copy())
    np.savez_compressed(OUT / "reference_outputs_color.npz", **g)
    print(f"wrote {len(g)} arrays, {(OUT / 'reference_outputs_color.npz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
