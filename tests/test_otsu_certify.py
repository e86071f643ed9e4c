"""The certificate behind the device Otsu scan (yam_hist.cu otsu_certify_kernel) is SOUND: whenever its NumPy
model certifies a histogram, the certified bin is the threshold of cv2's sequential fp64 recurrence
(libyamb200's host scan = oracle.otsu_from_hist, see test_host_helpers.py), and when it does not, the
recurrence's threshold is at or before the bin where the model lets the chain stop."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from oracle import np_oracle as O
from otsu_certify_model import certify
from yamimageprocessor_b200 import _lib, synth


@pytest.fixture(scope="module")
def chain_t():
    lib = _lib.load()

    def run(h):
        h = np.ascontiguousarray(h, dtype=np.uint64)
        out = C.c_int(0)
        _lib.check("yam_otsu_from_hist", lib.yam_otsu_from_hist(h.ctypes.data_as(C.c_void_p), int(h.size), C.byref(out)))
        return out.value

    return run


def _random_hist(rng, kind):
    x = np.arange(65536, dtype=np.float64)
    if kind == 0:    # two Gaussian modes
        m1, m2 = sorted(rng.uniform(500, 65000, 2)); s1, s2 = rng.uniform(50, 8000, 2); a1, a2 = rng.uniform(1e2, 1e6, 2)
        return (a1 * np.exp(-0.5 * ((x - m1) / s1) ** 2) + a2 * np.exp(-0.5 * ((x - m2) / s2) ** 2)).astype(np.int64)
    if kind == 1:    # symmetric: exact ties in the reals
        m = rng.integers(1000, 30000); s = rng.uniform(100, 3000); a = rng.uniform(1e2, 1e5)
        g = (a * np.exp(-0.5 * ((x - m) / s) ** 2)).astype(np.int64)
        return g + g[::-1]
    if kind == 2:    # sampled pixels, N not a power of two
        N = int(10 ** rng.uniform(3, 6.5)); m1, m2 = sorted(rng.uniform(500, 65000, 2)); s1, s2 = rng.uniform(50, 4000, 2)
        w = rng.uniform(0.05, 0.95)
        v = np.where(rng.random(N) < w, rng.normal(m1, s1, N), rng.normal(m2, s2, N))
        return np.bincount(np.clip(v, 0, 65535).astype(np.int64), minlength=65536)
    if kind == 3:    # a few spikes
        h = np.zeros(65536, np.int64); k = rng.integers(2, 12)
        h[rng.integers(0, 65536, k)] = rng.integers(1, 10 ** rng.integers(1, 9), k)
        return h
    if kind == 4:    # 8-bit
        return rng.integers(0, 10 ** rng.integers(1, 7), 256)
    lo = rng.integers(0, 60000); wd = int(rng.integers(2, 5000))   # narrow dense range
    h = np.zeros(65536, np.int64)
    h[lo:lo + wd] = rng.integers(0, 10 ** rng.integers(1, 6), min(wd, 65536 - lo))
    return h


def test_certificate_is_sound_on_random_histograms(chain_t):
    rng = np.random.default_rng(7)
    certified = 0
    cases = 0
    for it in range(360):
        h = _random_hist(rng, it % 6)
        if h.sum() == 0:
            continue
        cases += 1
        want = chain_t(h)
        ok, t, _, kmax = certify(h)
        if ok:
            certified += 1
            assert t == want, f"case {it}: certified {t}, recurrence {want}"
        else:
            assert want <= kmax, f"case {it}: recurrence {want} beyond kmax {kmax}"
    assert certified > cases // 4           # the certificate is not vacuous


def test_certificate_decides_microscopy_frames(chain_t):
    """Frames like the bench's (and their Gaussian / CLAHE outputs) are certified: no sequential scan."""
    for size, seed in ((512, 3), (1024, 5), (1000, 9)):
        x = synth.nuclei(size, size, seed)
        g = O.gaussian(x, 11, 0.0)
        for img in (x, g, O.clahe(g, 2.0, (8, 8))):
            h = O.histogram(img)
            ok, t, _, _ = certify(h)
            assert ok and t == chain_t(h) == O.otsu_from_hist(h)


def test_certificate_edge_cases(chain_t):
    z = np.zeros(65536, np.int64)
    assert certify(z) == (True, 0, 0, -1)
    one = z.copy(); one[100] = 5
    assert certify(one)[:2] == (True, 0) and chain_t(one) == 0              # nothing is ever evaluated
    two = z.copy(); two[100] = 5; two[60000] = 7
    ok, t, nc, kmax = certify(two)
    assert not ok and chain_t(two) <= kmax                                    # empty bins tie exactly
    big = z.copy(); big[65535] = 1 << 40; big[3] = 1 << 40
    assert certify(big)[0] is False                                          # first moment beyond 2^53
    eps_front = z.copy(); eps_front[10] = 2; eps_front[20000] = (1 << 24) - 4; eps_front[40000] = 2
    ok, t, _, kmax = certify(eps_front)                                      # q1 == FLT_EPSILON exactly at bin 10
    assert (ok and t == chain_t(eps_front)) or chain_t(eps_front) <= kmax


def test_equalised_images_have_comb_histograms_and_fall_back(chain_t):
    """A pure LUT-transformed image (equalizeHist, single-tile CLAHE) has a comb histogram: empty bins tie exactly,
    so the certificate must refuse -- and the bin where it lets the exact chain stop must still cover the answer."""
    rng = np.random.default_rng(3)
    v = np.clip(rng.normal(9000, 1200, 1 << 20), 0, 65535).astype(np.int64)
    v[: 1 << 17] = np.clip(rng.normal(30000, 2500, 1 << 17), 0, 65535).astype(np.int64)
    h = np.bincount(v, minlength=65536)
    lut = np.round(np.cumsum(h) * (65535.0 / h.sum())).astype(np.int64)          # equalisation LUT: steep where counts are dense
    comb = np.bincount(lut[v], minlength=65536)
    assert (comb == 0).sum() > 50000                                             # mostly empty bins
    ok, t, _, kmax = certify(comb)
    want = chain_t(comb)
    assert want == O.otsu_from_hist(comb)
    assert (ok and t == want) or (not ok and want <= kmax < 65535)
    twelve = np.bincount((v >> 4) << 4, minlength=65536)                         # 12-bit data in a 16-bit container
    ok, t, _, kmax = certify(twelve)
    assert not ok and chain_t(twelve) <= kmax
