"""SURVEY.md 8(f) N4 on the CPU: the oracle's watershed front half and chamfer distance transform against
the fixtures produced by the unmodified reference / its cv2 calls (tests/golden/make_golden_n4.py), and
against live cv2 when it is importable."""
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


@pytest.mark.parametrize("i", [0, 1, 2])
def test_watershed_front_half_matches_reference_intermediates(i):
    g = np.load(GOLD)
    k, oi, di, f = g[f"params_{i}"]
    gray = O.bgr2gray(g[f"in_bgr_{i}"])
    assert np.array_equal(gray, g[f"gray_{i}"])
    thresh, opening, sure_bg, dist, sure_fg, markers = O.watershed_front_half(gray, int(k), int(oi), int(di), float(f))
    assert np.array_equal(thresh, g[f"thresh_{i}"]) and np.array_equal(opening, g[f"opening_{i}"])
    assert np.array_equal(sure_bg, g[f"sure_bg_{i}"])
    assert np.array_equal(dist, g[f"dist_{i}"]), "float32 two-pass chamfer restatement differs from cv2.distanceTransform"
    assert np.array_equal(sure_fg, g[f"sure_fg_{i}"])
    assert np.array_equal(canon_markers(markers), canon_markers(g[f"markers_{i}"]))
    assert markers.max() >= 3     # the fixture has several markers, not just the background


def test_distance_transform_matches_live_cv2(rng):
    cv2 = pytest.importorskip("cv2")
    for shape, dens in (((40, 50), 0.8), ((33, 71), 0.95), ((64, 64), 0.6), ((50, 40), 0.995), ((20, 90), 1.0)):
        m = (rng.random(shape) < dens).astype(np.uint8) * 255
        m[3:18, 8:35] = 255          # a block with larger distances
        m[0, 0] = 0
        got, want = O.distance_transform_l2_5(m), cv2.distanceTransform(m, cv2.DIST_L2, 5)
        if want.max() < 40:
            assert np.array_equal(got, want), (shape, dens)
        else:
            # long paths: the wheel's scan (IPP) and the restatement round a few sums differently (one ulp,
            # ~1e-7 relative) -- float32 path sums are order dependent; tolerance 1e-6 relative
            assert (np.abs(got - want) <= 1e-6 * np.maximum(want, 1)).all(), (shape, dens)
            assert (got == want).mean() > 0.98


def test_region_shape_columns_known_answers():
    lab = np.zeros((12, 16), np.int32)
    lab[2:6, 3:11] = 1          # 4 x 8 rectangle: long axis along the columns
    lab[8:11, 1:2] = 2          # 3 x 1 vertical bar
    lab[8, 10] = 3              # single pixel
    c = O.region_shape_columns(lab)
    assert np.allclose(c["extent"], [1.0, 1.0, 1.0])
    # central moments of an a x b rectangle: (a^2 - 1) / 12 and (b^2 - 1) / 12
    l1, l2 = (8 * 8 - 1) / 12.0, (4 * 4 - 1) / 12.0
    assert np.isclose(c["eccentricity"][0], np.sqrt(1 - l2 / l1))
    assert np.isclose(abs(c["orientation"][0]), np.pi / 2) and np.isclose(c["orientation"][1], 0.0)
    assert c["eccentricity"][1] == 1.0 and c["eccentricity"][2] == 0.0
