"""Pin the NumPy oracle against the live third-party library the reference calls (cv2 4.13.0).
Skipped when cv2 is not importable; the committed golden fixtures (test_golden.py) pin it then."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import np_oracle as O

cv2 = pytest.importorskip("cv2")

U8, U16 = np.uint8, np.uint16


def rnd(rng, shape, dt):
    return rng.integers(0, (255 if dt == U8 else 65535) + 1, shape, dtype=dt)


def eq(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    assert int((a != b).sum()) == 0


@pytest.mark.parametrize("dt", [U8, U16])
def test_gray_and_gaussian_fixed(rng, dt):
    a = rnd(rng, (37, 53, 3), dt)
    eq(O.bgr2gray(a), cv2.cvtColor(a, cv2.COLOR_BGR2GRAY))
    for shp in ((64, 64), (33, 71), (5, 9)):
        for k in (1, 3, 5, 7, 9, 11, 13, 15, 17, 25):
            g = rnd(rng, shp, dt)
            eq(O.gaussian_fixed(g, k, 0), cv2.GaussianBlur(g, (k, k), 0))
        for s in (2.0, 1.3, 3.0):
            g = rnd(rng, shp, dt)
            eq(O.gaussian_fixed(g, O.ksize_from_sigma(s, dt == U8), s), cv2.GaussianBlur(g, (0, 0), s))


def test_gaussian_kernel_taps_quantise_like_cv2():
    for k in range(1, 32, 2):
        for s in (0, 0.5, 1.0, 2.0, 2.5, 3.3):
            ours, theirs = O.gaussian_kernel(k, s), cv2.getGaussianKernel(k, s).ravel()
            assert np.abs(ours - theirs).max() < 1e-15
            assert np.array_equal(ours.astype(np.float32), theirs.astype(np.float32))
            for bits in (8, 16):
                assert np.array_equal(O.fixed_kernel(ours, bits), O.fixed_kernel(theirs, bits))


def test_gaussian_f32_order(rng):
    for shp in ((64, 64), (128, 256), (40, 72)):  # W % 8 == 0 -> bit-exact
        for k in (1, 3, 5, 7, 9, 11, 21, 31):
            for border, cvb in (("reflect101", cv2.BORDER_REFLECT_101), ("replicate", cv2.BORDER_REPLICATE)):
                a = rng.integers(0, 65536, shp).astype(np.float32)
                eq(O.gaussian_f32(a, k, 0, border), cv2.GaussianBlur(a, (k, k), 0, borderType=cvb))
    a = rng.integers(0, 65536, (31, 100)).astype(np.float32)  # tail columns: <= 1e-5 relative
    o, c = O.gaussian_f32(a, 11, 0, "replicate"), cv2.GaussianBlur(a, (11, 11), 0, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(o[:, :96], c[:, :96])
    assert (np.abs(o - c) / np.maximum(np.abs(c), 1e-9)).max() <= 1e-5


@pytest.mark.parametrize("dt", [U8, U16])
def test_pointwise_and_filters(rng, dt):
    for shp in ((40, 50), (33, 71), (7, 9)):
        a = rnd(rng, shp, dt)
        for k in (3, 5):
            eq(O.median(a, k), cv2.medianBlur(a, k))
        for k in (3, 5, 7):
            eq(O.box(a, k), cv2.blur(a, (k, k)))
        b = np.maximum(a, dt(23))
        for al, be in ((0, 255), (10, 200), (255, 0)):
            eq(O.normalize_minmax(b, al, be), cv2.normalize(b, None, al, be, cv2.NORM_MINMAX))
        for al, be in ((1.0, 0), (1.5, -20), (0.37, 12.5), (2.9, 100)):
            eq(O.convert_scale_abs(a, al, be), cv2.convertScaleAbs(a, alpha=al, beta=be))
        eq(O.threshold_binary(a, 100.7, 255), cv2.threshold(a, 100.7, 255, cv2.THRESH_BINARY)[1])
        t, th = cv2.threshold(a, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        to, tho = O.otsu_threshold(a, 255)
        assert t == to
        eq(tho, th)
        for clip, grid in ((2.0, (8, 8)), (4.0, (4, 6)), (40, (3, 5)), (0, (8, 8))):
            eq(O.clahe(a, clip, grid), cv2.createCLAHE(clipLimit=clip, tileGridSize=grid).apply(a))


def test_u8_only_ops(rng):
    a = rnd(rng, (64, 80), U8)
    eq(O.equalize_hist(a), cv2.equalizeHist(a))
    eq(O.lut_u8(a, O.gamma_table(2.2)), cv2.LUT(a, O.gamma_table(2.2)))
    for shp in ((64, 64), (40, 72), (33, 71)):
        for blk, C in ((3, 2), (5, 2), (11, 2), (31, -3), (11, 2.5)):
            g = rnd(rng, shp, U8)
            eq(O.adaptive_threshold(g, blk, C),
               cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, blk, C))


@pytest.mark.parametrize("dt", [U8, U16])
def test_morphology(rng, dt):
    a = rnd(rng, (40, 50), dt)
    for name, cvs in (("Rectangular", cv2.MORPH_RECT), ("Elliptical", cv2.MORPH_ELLIPSE), ("Cross", cv2.MORPH_CROSS)):
        for k in (1, 2, 3, 4, 5, 7, 9, 15):
            se = cv2.getStructuringElement(cvs, (k, k))
            assert np.array_equal(O.structuring_element(name, k), se)
            for it in (1, 2):
                eq(O.erode(a, name, k, it), cv2.erode(a, se, iterations=it))
                eq(O.dilate(a, name, k, it), cv2.dilate(a, se, iterations=it))
                eq(O.morph_open(a, name, k, it), cv2.morphologyEx(a, cv2.MORPH_OPEN, se, iterations=it))
                eq(O.morph_close(a, name, k, it), cv2.morphologyEx(a, cv2.MORPH_CLOSE, se, iterations=it))


def test_ccl_and_region_props(rng):
    for dens in (0.1, 0.3, 0.5, 0.7):
        m = (rng.random((40, 60)) < dens).astype(U8) * 255
        n1, l1 = O.ccl_label(m)
        n2, l2 = O.ccl_label_python(m)
        assert n1 == n2 and np.array_equal(l1, l2)
        n4, l4, stats, cent = cv2.connectedComponentsWithStats(m, connectivity=8)
        assert n4 - 1 == n1
        l4c = O.canonicalise_labels(l4)
        assert np.array_equal(l1, l4c)
        inten = rnd(rng, (40, 60), U16)
        rp = O.region_props(l1, inten, n1)
        for lab in range(1, n4):
            ys, xs = np.nonzero(l4 == lab)
            i = l4c[ys[0], xs[0]] - 1
            assert rp["area"][i] == stats[lab, cv2.CC_STAT_AREA]
            assert tuple(rp["bbox"][i]) == (stats[lab, 1], stats[lab, 0], stats[lab, 1] + stats[lab, 3], stats[lab, 0] + stats[lab, 2])
            assert abs(rp["centroid_col"][i] - cent[lab, 0]) < 1e-9 and abs(rp["centroid_row"][i] - cent[lab, 1]) < 1e-9
            assert abs(rp["mean_intensity"][i] - inten[l4 == lab].mean()) < 1e-9


# ---------------------------------------------------------------------------------------------
# SURVEY.md 8(f) N3 steps
@pytest.mark.parametrize("dt", [U8, U16])
def test_add_weighted_and_sharpen(rng, dt):
    for shp in ((33, 71), (64, 64), (5, 9)):
        a, b = rnd(rng, shp, dt), rnd(rng, shp, dt)
        for al, be, ga in ((2.0, -1.0, 0.0), (3.3, -2.3, 0.0), (1.1, -0.1, 0.0), (0.25, 0.6, 7.5), (1.7, -0.7, -3.0)):
            eq(O.add_weighted(a, al, b, be, ga), cv2.addWeighted(a, al, b, be, ga))
        for s in (1.0, 0.35, 2.3, 5.0, 0.0):
            want = cv2.addWeighted(a, 1 + s, cv2.GaussianBlur(a, (0, 0), sigmaX=3), -s, 0)
            eq(O.sharpen(a, s), want)


@pytest.mark.parametrize("dt", [U8, U16])
def test_edge_operators(rng, dt):
    for shp in ((33, 71), (64, 64), (7, 5), (3, 3)):
        g = rnd(rng, shp, dt)
        for k in (1, 3, 5, 7):
            gx, gy = cv2.Sobel(g, cv2.CV_64F, 1, 0, ksize=k), cv2.Sobel(g, cv2.CV_64F, 0, 1, ksize=k)
            eq(O.sobel_magnitude(g, k), np.uint8(np.clip(cv2.magnitude(gx, gy), 0, 255)))
            eq(O.laplacian_abs(g, k), np.uint8(np.clip(np.abs(cv2.Laplacian(g, cv2.CV_64F, ksize=k)), 0, 255)))
        # low-amplitude input so that the clip at 255 does not hide the arithmetic
        small = (g >> (4 if dt == U8 else 12)).astype(dt)
        gx, gy = cv2.Sobel(small, cv2.CV_64F, 1, 0, ksize=3), cv2.Sobel(small, cv2.CV_64F, 0, 1, ksize=3)
        eq(O.sobel_magnitude(small, 3), np.uint8(np.clip(cv2.magnitude(gx, gy), 0, 255)))
        kx = np.array([[1, 0, -1], [1, 0, -1], [1, 0, -1]])
        fx, fy = cv2.filter2D(small, -1, kx), cv2.filter2D(small, -1, kx.T)
        want = np.uint8(np.clip(cv2.magnitude(fx.astype(np.float32), fy.astype(np.float32)), 0, 255))
        diff = O.prewitt_magnitude(small).astype(np.int16) - want.astype(np.int16)
        assert diff.min() >= 0 and diff.max() <= 1  # approximate float32 sqrt inside cv2.magnitude (see test_golden.py)


def test_moment_algebra_vs_cv2(rng):
    """host/moments.py (product-side float64 algebra on exact integer sums) against cv2.moments /
    cv2.HuMoments on random masks: raw moments exact while they fit 2^53, Hu moments to 1e-9."""
    from yamimageprocessor_b200.host import moments as M

    for shape in ((48, 64), (33, 71), (200, 300), (480, 640)):
        mask = ((rng.random(shape) < 0.3).astype(np.uint8)) * 255
        mask[shape[0] // 4: shape[0] // 2, shape[1] // 3: shape[1] // 2] = 255
        rows = np.zeros((shape[0], 4), np.int64)
        ys, xs = np.nonzero(mask)
        for p in range(4):
            np.add.at(rows[:, p], ys, xs.astype(np.int64) ** p)
        raw = M.raw_moments_from_rows(rows)
        ref = cv2.moments(mask)
        for key, val in raw.items():
            assert val == ref[key], key
        full = M.complete_moments(raw)
        for key in ("mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03", "nu20", "nu11", "nu02", "nu30", "nu21", "nu12", "nu03"):
            assert abs(full[key] - ref[key]) <= 1e-9 * max(1.0, abs(ref[key])), key
        np.testing.assert_allclose(M.hu_moments(full), cv2.HuMoments(ref).ravel(), rtol=1e-9, atol=0)
    empty = M.complete_moments(M.raw_moments_from_rows(np.zeros((8, 4), np.int64)))
    assert all(v == 0.0 for v in empty.values())


def test_histogram_statistics_vs_reference_formula(rng):
    from scipy.stats import kurtosis, skew

    from yamimageprocessor_b200.host import moments as M

    for img in (rnd(rng, (64, 80), U8), (rnd(rng, (90, 70), U8) >> 2).astype(U8)):
        hist = cv2.calcHist([img], [0], None, [256], [0, 256]).flatten()
        total = np.sum(hist)
        pixels = np.arange(256)
        mean_val = np.sum(pixels * hist) / total
        var_val = np.sum(((pixels - mean_val) ** 2) * hist) / total
        data = np.repeat(pixels, hist.astype(int))
        want = [mean_val, var_val, skew(data), kurtosis(data)]
        st = M.histogram_statistics(np.bincount(img.ravel(), minlength=256))
        np.testing.assert_allclose([st["mean"], st["variance"], st["skewness"], st["kurtosis"]], want, rtol=1e-9, atol=0)
