"""cv2.moments / cv2.HuMoments algebra on exact row power sums (hu_moments_data,
``core/extraction.py:100-105``) and the histogram statistics of ``histogram_data`` (``:280-290``).

The device delivers exact integers (per-row power sums of the mask, 256-bin counts); the handful
of float64 operations that turn them into the reference's table run here.  cv2's build contracts
some of the central-moment expressions into FMAs, so Hu moments agree to ~1e-13 relative, not bit
for bit (tolerance 1e-9 in the tests).
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def raw_moments_from_rows(rows: np.ndarray, weight: int = 255) -> Dict[str, float]:
    """``rows[y] = (count, sum x, sum x^2, sum x^3)`` -> m00..m03 as float64 of the EXACT integer sums
    (cv2.moments on a {0,255} image weighs every set pixel with 255)."""
    rows = np.asarray(rows)
    ys = [int(v) for v in range(rows.shape[0])]
    cols = [[int(v) for v in rows[:, p]] for p in range(4)]
    out = {}
    for p in range(4):
        for q in range(4 - p):
            total = 0
            for y, s in zip(ys, cols[p]):
                if s:
                    total += (y ** q) * s
            out[f"m{p}{q}"] = float(weight * total)
    return out


def complete_moments(m: Dict[str, float]) -> Dict[str, float]:
    """Central and normalised central moments (cv2 completeMomentState)."""
    m00, m10, m01 = m["m00"], m["m10"], m["m01"]
    cx = cy = inv = 0.0
    if abs(m00) > np.finfo(np.float64).eps:
        inv = 1.0 / m00
        cx, cy = m10 * inv, m01 * inv
    mu20 = m["m20"] - m10 * cx
    mu11 = m["m11"] - m10 * cy
    mu02 = m["m02"] - m01 * cy
    mu30 = m["m30"] - cx * (3 * mu20 + cx * m10)
    mu21 = m["m21"] - cx * (2 * mu11 + cx * m01) - cy * mu20
    mu12 = m["m12"] - cy * (2 * mu11 + cy * m10) - cx * mu02
    mu03 = m["m03"] - cy * (3 * mu02 + cy * m01)
    s2 = inv * inv
    s3 = s2 * float(np.sqrt(abs(inv)))
    out = dict(m)
    out.update(mu20=mu20, mu11=mu11, mu02=mu02, mu30=mu30, mu21=mu21, mu12=mu12, mu03=mu03,
               nu20=mu20 * s2, nu11=mu11 * s2, nu02=mu02 * s2,
               nu30=mu30 * s3, nu21=mu21 * s3, nu12=mu12 * s3, nu03=mu03 * s3)
    return out


def hu_moments(m: Dict[str, float]) -> np.ndarray:
    """cv2.HuMoments on the normalised central moments."""
    t0, t1 = m["nu30"] + m["nu12"], m["nu21"] + m["nu03"]
    q0, q1 = t0 * t0, t1 * t1
    n4 = 4 * m["nu11"]
    s, d = m["nu20"] + m["nu02"], m["nu20"] - m["nu02"]
    hu = np.zeros(7, np.float64)
    hu[0] = s
    hu[1] = d * d + n4 * m["nu11"]
    hu[3] = q0 + q1
    hu[5] = d * (q0 - q1) + n4 * t0 * t1
    t0 *= q0 - 3 * q1
    t1 *= 3 * q0 - q1
    q0, q1 = m["nu30"] - 3 * m["nu12"], 3 * m["nu21"] - m["nu03"]
    hu[2] = q0 * q0 + q1 * q1
    hu[4] = q0 * t0 + q1 * t1
    hu[6] = q1 * t0 - q0 * t1
    return hu


def histogram_statistics(counts: np.ndarray) -> Dict[str, float]:
    """mean / variance / skewness / kurtosis of ``histogram_data`` from the 256 bin counts.

    The reference feeds ``cv2.calcHist`` (float32 counts) into float64 sums and scipy's biased
    skew / Fisher kurtosis of the repeated data; the same central moments are formed here from the
    counts (agreement ~1e-12 relative; skewness / kurtosis are nan for a constant image, like scipy)."""
    hist = np.asarray(counts, dtype=np.float64).astype(np.float32)   # calcHist returns float32 counts
    total = np.sum(hist) if np.sum(hist) != 0 else 1
    pixels = np.arange(256)
    mean_val = np.sum(pixels * hist) / total
    variance_val = np.sum(((pixels - mean_val) ** 2) * hist) / total
    n = hist.astype(int).astype(np.float64)
    nn = n.sum()
    if nn == 0:
        return {"mean": float(mean_val), "variance": float(variance_val), "skewness": 0.0, "kurtosis": 0.0}
    mu = np.sum(pixels * n) / nn
    dev = pixels - mu
    m2, m3, m4 = np.sum(n * dev ** 2) / nn, np.sum(n * dev ** 3) / nn, np.sum(n * dev ** 4) / nn
    with np.errstate(divide="ignore", invalid="ignore"):
        skew = m3 / m2 ** 1.5 if m2 > 0 else float("nan")
        kurt = m4 / m2 ** 2 - 3.0 if m2 > 0 else float("nan")
    return {"mean": float(mean_val), "variance": float(variance_val), "skewness": float(skew), "kurtosis": float(kurt)}


__all__ = ["complete_moments", "histogram_statistics", "hu_moments", "raw_moments_from_rows"]
