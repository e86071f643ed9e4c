"""Oracle vs fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "reference_outputs.npz")


@pytest.fixture(scope="module")
def meta():
    return json.loads((GOLD / "reference_meta.json").read_text())


def eq(a, b, what=""):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    assert int((a != b).sum()) == 0, what


@pytest.mark.parametrize("tag", ["u8", "u16"])
def test_preprocessing_modules(gold, tag):
    bgr, gray, noise = gold[f"in_bgr_{tag}"], gold[f"in_gray_{tag}"], gold[f"in_noise_{tag}"]
    dt = gray.dtype.type
    eq(O.bgr2gray(bgr), gold[f"grayscale_{tag}"], "Grayscale")
    for k in (3, 5, 11, 15):
        eq(O.gaussian(noise, k, 0.0), gold[f"gauss{k}_{tag}"], f"NoiseReduction Gaussian {k}")
    for k in (3, 5):
        eq(O.median(noise, k), gold[f"median{k}_{tag}"], f"NoiseReduction Median {k}")
    eq(O.normalize_minmax(np.maximum(gray, dt(9)), 0, 255), gold[f"normalize_{tag}"], "IntensityNormalization")
    eq(O.normalize_minmax(np.maximum(gray, dt(9)), 10, 200), gold[f"normalize_10_200_{tag}"], "IntensityNormalization 10..200")
    eq(O.convert_scale_abs(noise, 1.5, -20), gold[f"brightness_{tag}"], "BrightnessContrast")
    eq(O.clahe(gray, 2.0, (8, 8)), gold[f"clahe_{tag}"], "CLAHE")
    eq(O.clahe(gray, 4.0, (3, 5)), gold[f"clahe_4_3x5_{tag}"], "CLAHE 3x5")
    eq(O.box(noise, 5), gold[f"box5_{tag}"], "box")


@pytest.mark.parametrize("tag", ["u8", "u16"])
def test_segmentation_functions(gold, tag):
    bgr, gray, noise = gold[f"in_bgr_{tag}"], gold[f"in_gray_{tag}"], gold[f"in_noise_{tag}"]
    eq(O.otsu_threshold(gray, 255)[1], gold[f"otsu_{tag}"], "otsu")
    eq(O.otsu_threshold(O.bgr2gray(bgr), 255)[1], gold[f"otsu_bgr_{tag}"], "otsu on colour")
    eq(O.threshold_binary(gray, 100, 255), gold[f"global100_{tag}"], "global")
    for shape in ("Rectangular", "Elliptical", "Cross"):
        for k, it in ((3, 1), (5, 2)):
            key = f"{shape.lower()}_{k}_{it}_{tag}"
            eq(O.morph_open(noise, shape, k, it), gold[f"open_{key}"], f"open {key}")
            eq(O.morph_close(noise, shape, k, it), gold[f"close_{key}"], f"close {key}")
            eq(O.dilate(noise, shape, k, it), gold[f"dilate_{key}"], f"dilate {key}")
            eq(O.erode(noise, shape, k, it), gold[f"erode_{key}"], f"erode {key}")


def test_u8_only(gold):
    eq(O.lut_u8(gold["in_noise_u8"], O.gamma_table(2.2)), gold["gamma22_u8"], "Gamma")
    eq(O.adaptive_threshold(gold["in_gray_u8"], 11, 2), gold["adaptive_11_2_u8"], "Adaptive 11,2")
    eq(O.adaptive_threshold(gold["in_gray_u8"], 31, -3), gold["adaptive_31_m3_u8"], "Adaptive 31,-3")
    eq(O.adaptive_threshold(O.bgr2gray(gold["in_bgr_u8"]), 11, 2), gold["adaptive_bgr_u8"], "Adaptive colour")
    eq(O.equalize_hist(gold["in_gray_u8"]), gold["equalize_u8"], "equalizeHist")


def test_connected_components(gold, meta):
    n, lab = O.ccl_label(gold["ccl_mask_u8"])
    assert n == meta["ccl_count"]
    eq(lab, O.canonicalise_labels(gold["ccl_labels_cv2"]), "labels (canonical)")
    props = O.region_props(lab, gold["in_gray_u8"], n)
    stats, cent, cvlab = gold["ccl_stats_cv2"], gold["ccl_centroids_cv2"], gold["ccl_labels_cv2"]
    canon = O.canonicalise_labels(cvlab)
    for l in range(1, stats.shape[0]):
        ys, xs = np.nonzero(cvlab == l)
        i = canon[ys[0], xs[0]] - 1
        assert props["area"][i] == stats[l, 4]
        assert tuple(props["bbox"][i]) == (stats[l, 1], stats[l, 0], stats[l, 1] + stats[l, 3], stats[l, 0] + stats[l, 2])
        assert abs(props["centroid_col"][i] - cent[l, 0]) < 1e-9 and abs(props["centroid_row"][i] - cent[l, 1]) < 1e-9


# ---------------------------------------------------------------------------------------------
# fixtures from tests/golden/make_golden_n3.py (unmodified reference)
@pytest.fixture(scope="module")
def gold_n3():
    return np.load(GOLD / "reference_outputs_n3.npz")


@pytest.mark.parametrize("tag", ["u8", "u16"])
def test_n3_steps(gold_n3, tag):
    g = gold_n3
    noise, ramp, bgr = g[f"in_noise_{tag}"], g[f"in_ramp_{tag}"], g[f"in_bgr_{tag}"]
    for s in (1.0, 0.35, 2.3):
        eq(O.sharpen(noise, s), g[f"sharpen_{s}_{tag}"], f"Sharpen {s}")
        eq(O.sharpen(ramp, s), g[f"sharpen_ramp_{s}_{tag}"], f"Sharpen ramp {s}")
    for ch in ("R", "G", "B", "All"):
        eq(O.select_channel(bgr, ch), g[f"select_{ch}_{tag}"], f"SelectChannel {ch}")
    eq(O.select_channel(noise, "All"), g[f"select_gray_All_{tag}"], "SelectChannel gray All")
    eq(O.select_channel(noise, "G"), g[f"select_gray_G_{tag}"], "SelectChannel gray G")
    for k in (1, 3, 5, 7):
        for name, img in (("noise", noise), ("ramp", ramp)):
            eq(O.sobel_magnitude(img, k), g[f"sobel{k}_{name}_{tag}"], f"Sobel {k} {name}")
            eq(O.laplacian_abs(img, k), g[f"laplacian{k}_{name}_{tag}"], f"Laplacian {k} {name}")
    eq(O.sobel_magnitude(O.bgr2gray(bgr), 3), g[f"sobel3_bgr_{tag}"], "Sobel colour")
    for bd in (0, 1, 5, 22, 23, 40):
        eq(O.remove_border_regions(noise, bd), g[f"border{bd}_{tag}"], f"Border Removal {bd}")
    eq(O.remove_border_regions(bgr, 5), g[f"border5_bgr_{tag}"], "Border Removal colour")
    # Prewitt: cv2.magnitude(float32) goes through an approximate square root in the wheel's IPP
    # build (cv2.magnitude(123, 0) == 122.99999), so exact squares can truncate one lower.  The
    # restatement uses the correctly rounded root; tolerance = north_star's 1 LSB on uint8 output.
    for name, img in (("noise", noise), ("ramp", ramp)):
        got, want = O.prewitt_magnitude(img), g[f"prewitt_{name}_{tag}"]
        diff = got.astype(np.int16) - want.astype(np.int16)
        assert np.abs(diff).max() <= 1, f"Prewitt {name}"   # tolerance: 1 LSB either way (SIMD-path dependent sqrt)
        assert (diff != 0).mean() < 0.01


def test_n3_channel_means(gold_n3):
    for ch in ("RG", "GB", "BR"):
        eq(O.select_channel(gold_n3["in_bgr_u8"], ch), gold_n3[f"select_{ch}_u8"], f"SelectChannel {ch}")
    with pytest.raises(TypeError):
        O.select_channel(gold_n3["in_bgr_u16"], "RG")


def _row_power_sums(mask):
    rows = np.zeros((mask.shape[0], 4), np.int64)
    ys, xs = np.nonzero(mask)
    xs = xs.astype(np.int64)
    for p in range(4):
        np.add.at(rows[:, p], ys, xs ** p)
    return rows


def test_n3_extraction_tables(gold_n3):
    """hu_moments_data / histogram_data of the reference vs the host algebra on exact integer inputs.
    Tolerance 1e-9 relative: cv2's build contracts some central-moment expressions into FMAs and
    scipy sums the repeated data pairwise, so the last bits differ (see host/moments.py)."""
    from yamimageprocessor_b200.host import moments as M

    for name in ("blob_u8", "noise_u8", "bgr_u8", "blob_u16"):
        img = gold_n3[f"in_{name}"]
        gray = O.bgr2gray(img) if img.ndim == 3 else img
        mask = O.otsu_threshold(gray, 255)[1]
        hu = M.hu_moments(M.complete_moments(M.raw_moments_from_rows(_row_power_sums(mask))))
        np.testing.assert_allclose(hu, gold_n3[f"hu_{name}"], rtol=1e-9, atol=0)
        if name.endswith("u8"):
            st = M.histogram_statistics(np.bincount(gray.ravel(), minlength=256))
            got = np.array([st["mean"], st["variance"], st["skewness"], st["kurtosis"]])
            np.testing.assert_allclose(got, gold_n3[f"histstats_{name}"], rtol=1e-9, atol=0)
