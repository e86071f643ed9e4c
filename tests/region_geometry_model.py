"""NumPy / pure-Python model of the device algorithms behind the ``perimeter`` and ``solidity`` columns
(``yam_region_perimeter``, ``yam_region_convex_area`` in ``csrc/yam_regiongeom.cu``): the same integer
steps the kernels take, so the arithmetic can be checked on the CPU against the oracle's literal
restatement of skimage (``oracle/np_oracle.py: perimeter4, convex_area``).  Test infrastructure only.

Perimeter: on the LABEL image.  A pixel of label L is a border pixel when one of its four neighbours
is not L (outside the image counts as not L).  A border pixel p is classed by
v = 1 + 2 * #(4-neighbours q: label L and border) + 10 * #(diagonal neighbours q: label L and border);
the kernel adds to one of three integer counters per label (class weight 1, sqrt 2, (1 + sqrt 2) / 2).

Convex area: skimage's hull is the hull of the edge midpoints of the region's pixels.  In doubled
coordinates (R, C) = (2r, 2c) those are (2r - 1, 2c), (2r + 1, 2c), (2r, 2c - 1), (2r, 2c + 1).  Index the
region's doubled rows by k = R - (2 r_min - 1), k = 0 .. 2H.  Per k only the smallest C (left chain) and the
largest C (right chain) can be hull vertices:
    k odd  (R = 2r)      V_left = 2 cmin(r) - 1,                  V_right = -(2 cmax(r) + 1)
    k even (R = 2r + 1)  V_left = 2 min(cmin(r), cmin(r + 1)),    V_right = -2 max(cmax(r), cmax(r + 1))
(the right chain is stored negated, so both are LOWER convex hulls of (k, V_k)).  One monotone-chain
pass per side; then, for every pixel row (odd k) the hull edge (ka, Va) - (kb, Vb) spanning it gives the
rational bound v = Va + (Vb - Va)(k - ka) / (kb - ka) and the side adds  side - ceil(v / 2)  to the
region's count (side = 0 left, 1 right): sum over rows of  floor(hi / 2) - ceil(lo / 2) + 1.
"""
from __future__ import annotations

import math

import numpy as np

MISSING = 0x7F7F7F7F     # cudaMemset(0x7f) pattern: "no pixel contributed to this k"

# class of a border pixel by v = 1 + 2 a + 10 d  (0 = weighs nothing)
PERIMETER_CLASS = np.zeros(50, np.int64)
PERIMETER_CLASS[[5, 7, 15, 17, 25, 27]] = 1
PERIMETER_CLASS[[21, 33]] = 2
PERIMETER_CLASS[[13, 23]] = 3
PERIMETER_WEIGHTS = np.array([0.0, 1.0, math.sqrt(2.0), (1.0 + math.sqrt(2.0)) / 2.0])


def perimeter_counts(labels: np.ndarray, n: int) -> np.ndarray:
    """int64 [n, 3]: border pixels of class 1, 2, 3 per label."""
    h, w = labels.shape
    lab = np.zeros((h + 4, w + 4), np.int64)          # two rings of label 0 = outside the image
    lab[2:-2, 2:-2] = labels

    def win(a, dy, dx, ring):
        """the image plus `ring` pixels around it, shifted by (dy, dx)"""
        o = 2 - ring
        return a[o + dy:o + dy + h + 2 * ring, o + dx:o + dx + w + 2 * ring]

    c1 = win(lab, 0, 0, 1)
    differs = np.zeros(c1.shape, bool)
    for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        differs |= win(lab, dy, dx, 1) != c1
    border = np.zeros(lab.shape, bool)
    border[1:-1, 1:-1] = (c1 > 0) & differs            # ring pixels are label 0: never border
    p, pb = win(lab, 0, 0, 0), win(border, 0, 0, 0)
    a = np.zeros(p.shape, np.int64)
    d = np.zeros(p.shape, np.int64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy == 0 and dx == 0:
                continue
            hit = (win(lab, dy, dx, 0) == p) & win(border, dy, dx, 0)
            if dy == 0 or dx == 0:
                a += hit
            else:
                d += hit
    cls = np.where(pb, PERIMETER_CLASS[1 + 2 * a + 10 * d], 0)
    out = np.zeros((n, 3), np.int64)
    for k in (1, 2, 3):
        sel = (cls == k) & (p > 0) & (p <= n)
        out[:, k - 1] = np.bincount(p[sel] - 1, minlength=n)[:n]
    return out


def perimeter_from_counts(counts: np.ndarray) -> np.ndarray:
    return counts[:, 0] * PERIMETER_WEIGHTS[1] + counts[:, 1] * PERIMETER_WEIGHTS[2] + counts[:, 2] * PERIMETER_WEIGHTS[3]


def _ceil_div(a: int, b: int) -> int:
    """ceil(a / b) for b > 0 the way the kernel computes it (C division truncates towards zero)."""
    return (a + b - 1) // b if a >= 0 else -((-a) // b)


def chain_points(labels: np.ndarray, n: int, props: np.ndarray):
    """The two V arrays (left, right-negated) of every label, as the fill kernel leaves them."""
    heights = np.where(props[:, 0] > 0, props[:, 6] - props[:, 4], 0)
    npts = np.where(props[:, 0] > 0, 2 * heights + 1, 0)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(npts, out=off[1:])
    vl = np.full(int(off[-1]), MISSING, np.int64)
    vr = np.full(int(off[-1]), MISSING, np.int64)
    ys, xs = np.nonzero(labels)
    for y, x in zip(ys.tolist(), xs.tolist()):
        lab = int(labels[y, x])
        if lab > n:
            continue
        base = int(off[lab - 1]) + 2 * (y - int(props[lab - 1, 4]))
        for dk, cl, cr in ((0, 2 * x, -2 * x), (1, 2 * x - 1, -(2 * x + 1)), (2, 2 * x, -2 * x)):
            vl[base + dk] = min(vl[base + dk], cl)
            vr[base + dk] = min(vr[base + dk], cr)
    return off, vl, vr


def chain_contribution(v: np.ndarray, side: int) -> int:
    """One (region, side) thread: lower hull of (k, v[k]) by a monotone chain, then the row sums."""
    stack = []
    for k in range(len(v)):
        vk = int(v[k])
        if vk == MISSING:
            continue
        while len(stack) >= 2:
            ka, kb = stack[-2], stack[-1]
            cross = (kb - ka) * (vk - int(v[ka])) - (int(v[kb]) - int(v[ka])) * (k - ka)
            if cross <= 0:
                stack.pop()
            else:
                break
        stack.append(k)
    total = 0
    for ka, kb in zip(stack[:-1], stack[1:]):
        va, vb = int(v[ka]), int(v[kb])
        den = kb - ka
        k = ka if ka & 1 else ka + 1
        while k < kb:                       # odd k = pixel rows, each owned by the edge with ka <= k < kb
            num = va * den + (vb - va) * (k - ka)
            total += side - _ceil_div(num, 2 * den)
            k += 2
    return total


def convex_area_model(labels: np.ndarray, n: int, props: np.ndarray) -> np.ndarray:
    off, vl, vr = chain_points(labels, n, props)
    out = np.zeros(n, np.int64)
    for i in range(n):
        a, b = int(off[i]), int(off[i + 1])
        if b > a:
            out[i] = chain_contribution(vl[a:b], 0) + chain_contribution(vr[a:b], 1)
    return out
