"""SURVEY.md 8(f) N4 on the GPU: watershed front half (inverse Otsu -> open -> dilate -> chamfer distance
transform -> markers) and the second-moment region columns, against the reference fixtures, the oracle
and live cv2.

Tolerance (stated): the distance transform is a float32 filter -> <= 1e-5 relative (north_star).  cv2's
sequential two-pass scan and the device's least-fixed-point relaxation differ by at most isolated ulps;
everything derived from integers is bit-exact."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as O
GOLD = Path(__file__).resolve().parent / "golden" / "reference_outputs_n4.npz"


def canon_markers(m):
    """markers: 0 unknown, 1 background, >= 2 components -> component ids renumbered raster-first"""
    comp = np.where(m > 1, m - 1, 0)
    return np.where(m > 1, O.canonicalise_labels(comp) + 1, m)

pytestmark = pytest.mark.gpu
DT_RTOL = 1e-5


def check_dist(got, want, what):
    assert got.dtype == np.float32 and got.shape == want.shape
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(want, 1e-30)
    rel[want == 0] = np.where(got[want == 0] == 0, 0.0, np.inf)
    assert rel.max() <= DT_RTOL, f"{what}: distance transform off by {rel.max():.3g} relative"
    return float((got == want).mean())


@pytest.mark.parametrize("i", [0, 1, 2])
def test_watershed_markers_match_reference_fixture(backend, i):
    g = np.load(GOLD)
    k, oi, di, f = g[f"params_{i}"]
    from yamimageprocessor_b200.modules import b200_backend as plugin

    out = plugin.watershed_markers_data(g[f"in_bgr_{i}"], int(k), int(oi), int(di), float(f))
    for name in ("thresh", "opening", "sure_bg"):
        assert np.array_equal(out[name], g[f"{name}_{i}"]), name
    exact = check_dist(out["dist"], g[f"dist_{i}"], f"fixture {i}")
    assert exact >= 0.999
    # sure_fg can only differ where the distance sits within the tolerance of the threshold
    thr = np.float32(f) * np.float32(g[f"dist_{i}"].max())
    differs = out["sure_fg"] != g[f"sure_fg_{i}"]
    assert not differs[np.abs(g[f"dist_{i}"] - thr) > DT_RTOL * thr].any()
    if not differs.any():
        assert np.array_equal(canon_markers(out["markers"]), canon_markers(g[f"markers_{i}"]))
        assert int(out["n_markers"][0]) == int(g[f"markers_{i}"].max()) - 1


def test_distance_transform_vs_oracle_and_cv2(backend, rng):
    cases = []
    for shape, dens in (((40, 50), 0.8), ((33, 71), 0.95), ((64, 64), 0.6), ((97, 130), 0.97)):
        m = (rng.random(shape) < dens).astype(np.uint8) * 255
        m[3:30, 8:35] = 255
        m[0, 0] = 0
        cases.append(m)
    for m in cases:
        got = backend.to_host(backend.distance_transform(backend.to_device(m)))
        check_dist(got, O.distance_transform_l2_5(m), f"oracle {m.shape}")
    cv2 = pytest.importorskip("cv2")
    # large distances: convergence needs several relaxation launches
    big = np.full((700, 900), 255, np.uint8)
    big[350, 10] = 0
    big[20:40, 600:650] = 0
    stack = np.stack([big, np.roll(big, 77, axis=1)])
    got = backend.to_host(backend.distance_transform(backend.to_device(stack)))
    assert backend.last_distance_launches > 3
    for fr in range(2):
        exact = check_dist(got[fr], cv2.distanceTransform(stack[fr], cv2.DIST_L2, 5), f"cv2 large {fr}")
        assert exact > 0.95      # distances up to ~900 px: one-ulp differences on ~1 % of the pixels
    # a nuclei-like frame at 2048^2
    from yamimageprocessor_b200 import synth

    fr = (synth.nuclei(2048, 2048, seed=9) >> 8).astype(np.uint8)
    mask = np.where(fr > 40, 255, 0).astype(np.uint8)
    got = backend.to_host(backend.distance_transform(backend.to_device(mask)))
    check_dist(got, cv2.distanceTransform(mask, cv2.DIST_L2, 5), "cv2 nuclei 2048")


def test_threshold_inv(backend, rng):
    for dt in (np.uint8, np.uint16):
        hi = 255 if dt == np.uint8 else 65535
        a = rng.integers(0, hi + 1, (3, 37, 53), dtype=dt)
        x = backend.to_device(a)
        t, _ = backend.otsu_threshold(x, want_image=False)
        got = backend.to_host(backend.threshold_inv(x, t_dev=t))
        for fr in range(3):
            assert np.array_equal(got[fr], O.threshold_binary_inv(a[fr], O.otsu_value(a[fr]), 255))
        assert np.array_equal(backend.to_host(backend.threshold_inv(x, thresh=100.7, maxval=200)), O.threshold_binary_inv(a, 100.7, 200))


def test_region_shape_columns(backend, rng):
    from yamimageprocessor_b200 import synth
    from yamimageprocessor_b200.backend import shape_columns
    from yamimageprocessor_b200.host.steps import region_table

    frame = synth.nuclei(600, 520, seed=4)
    m = O.morph_close(O.morph_open(O.adaptive_threshold(frame, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1)
    m[100:104, 50:400] = 255          # an elongated horizontal region
    m[200:500, 300:302] = 255         # an elongated vertical region
    n, lab = O.ccl_label(m)
    labels = backend.to_device(lab)
    mom = backend.to_host(backend.region_moments(labels, n))
    ys, xs = np.nonzero(lab)
    l = lab[ys, xs].astype(np.int64) - 1
    for col, wts in enumerate((ys.astype(np.int64) ** 2, xs.astype(np.int64) ** 2, ys.astype(np.int64) * xs)):
        want = np.zeros(n, np.int64)
        np.add.at(want, l, wts)
        assert np.array_equal(mom[:, col], want), f"second-order sum {col}"
    table = region_table(backend, labels, backend.to_device(frame), n)
    want = O.region_shape_columns(lab, n)
    for name in ("extent", "eccentricity", "orientation"):
        # tolerance: float64 on both sides, different summation (raw vs centred moments)
        assert np.allclose(table[name], want[name], rtol=1e-7, atol=1e-7), name
    assert {"area", "centroid", "bbox", "mean_intensity", "extent", "eccentricity", "orientation"} <= set(table)
