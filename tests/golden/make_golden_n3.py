#!/usr/bin/env python
"""Golden vectors for the SURVEY.md 8(f) N3 steps (Sharpen, SelectChannel, Sobel / Prewitt /
Laplacian, Border Removal), produced by the UNMODIFIED reference from /root/reference.

    python tests/golden/make_golden_n3.py      # build container only (needs the reference + cv2 4.13.0)

Same stubbing as make_golden.py; inputs are stored next to the outputs in reference_outputs_n3.npz.
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
    from modules import preprocessing as mp
    from core import segmentation as cs

    rng = np.random.default_rng(20261019)
    H, W = 45, 70
    g: dict[str, np.ndarray] = {}
    for tag, dt in (("u8", np.uint8), ("u16", np.uint16)):
        hi = 255 if dt == np.uint8 else 65535
        noise = rng.integers(0, hi + 1, (H, W), dtype=dt)
        yy, xx = np.mgrid[0:H, 0:W]
        ramp = np.clip((np.sin(yy / 5.0) * np.cos(xx / 7.0) * 0.4 + 0.5) * hi + rng.normal(0, hi * 0.02, (H, W)), 0, hi).astype(dt)
        bgr = rng.integers(0, hi + 1, (H, W, 3), dtype=dt)
        g[f"in_noise_{tag}"], g[f"in_ramp_{tag}"], g[f"in_bgr_{tag}"] = noise, ramp, bgr
        for s in (1.0, 0.35, 2.3):
            g[f"sharpen_{s}_{tag}"] = mp.SharpenModule().process(noise, strength=s)
            g[f"sharpen_ramp_{s}_{tag}"] = mp.SharpenModule().process(ramp, strength=s)
        for ch in ("R", "G", "B", "All"):
            g[f"select_{ch}_{tag}"] = mp.SelectChannelModule().process(bgr, channel=ch)
        g[f"select_gray_All_{tag}"] = mp.SelectChannelModule().process(noise, channel="All")
        g[f"select_gray_G_{tag}"] = mp.SelectChannelModule().process(noise, channel="G")
        for k in (1, 3, 5, 7):
            for name, img in (("noise", noise), ("ramp", ramp)):
                g[f"sobel{k}_{name}_{tag}"] = cs.sobel_operator(img, k)
                g[f"laplacian{k}_{name}_{tag}"] = cs.laplacian_operator(img, k)
        g[f"sobel3_bgr_{tag}"] = cs.sobel_operator(bgr, 3)
        for name, img in (("noise", noise), ("ramp", ramp)):
            try:
                g[f"prewitt_{name}_{tag}"] = cs.prewitt_operator(img)
            except Exception as exc:  # noqa: BLE001 - recorded, not hidden
                print(f"prewitt {name} {tag}: {type(exc).__name__}: {exc}")
        for bd in (0, 1, 5, 22, 23, 40):
            g[f"border{bd}_{tag}"] = cs.remove_border_regions(noise, bd)
        g[f"border5_bgr_{tag}"] = cs.remove_border_regions(bgr, 5)
    for ch in ("RG", "GB", "BR"):
        g[f"select_{ch}_u8"] = mp.SelectChannelModule().process(g["in_bgr_u8"], channel=ch)
    # extraction tables (core/extraction.py:100-105, 280-290); blob image so that Otsu splits it
    from core import extraction as ce

    yy, xx = np.mgrid[0:96, 0:128]
    blob = np.zeros((96, 128))
    for _ in range(7):
        cy, cx, r = rng.integers(10, 86), rng.integers(10, 118), rng.integers(4, 12)
        blob += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * r * r))
    blob_u8 = np.clip(blob * 200 + rng.normal(8, 3, blob.shape), 0, 255).astype(np.uint8)
    g["in_blob_u8"] = blob_u8
    for name, img in (("blob", blob_u8), ("noise", g["in_noise_u8"]), ("bgr", g["in_bgr_u8"])):
        g[f"hu_{name}_u8"] = ce.hu_moments_data(img).to_numpy(dtype=np.float64).ravel()
        g[f"histstats_{name}_u8"] = ce.histogram_data(img).to_numpy(dtype=np.float64).ravel()
    blob_u16 = (blob_u8.astype(np.uint16) << 8) | 17
    g["in_blob_u16"] = blob_u16
    g["hu_blob_u16"] = ce.hu_moments_data(blob_u16).to_numpy(dtype=np.float64).ravel()
    np.savez_compressed(OUT / "reference_outputs_n3.npz", **g)
    print(f"wrote {len(g)} arrays, {(OUT / 'reference_outputs_n3.npz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
