"""NumPy model of ``otsu_certify_kernel`` (yamimageprocessor_b200/csrc/yam_hist.cu): the same interval bounds,
evaluated with NumPy's IEEE doubles.  Test infrastructure: the CPU suite checks that whatever this model
certifies IS the threshold of the sequential recurrence (cv2 getThreshVal_Otsu_16u restated in
oracle/np_oracle.py:otsu_from_hist / libyamb200's host scan), and the GPU suite checks the kernel against it.

    certify(hist) -> (certified, t, competitors, kmax)
      certified  the certificate decides the frame; t is then the threshold
      kmax       otherwise: the last bin that can still win (where the exact chain may stop)
"""
from __future__ import annotations

import numpy as np

U = 2.0 ** -53
EPS = float(np.finfo(np.float32).eps)
INFL = 1.01


def differential_excludes(ks, i, q1, q2, mu1, mu2, T, M, rq, rT, nonc, pow2):
    """True where bin k provably loses against bin i although their sigma intervals overlap.

    With A = T - M q (= q q2 (mu1 - mu2)) the recurrence's value is F(q^, T^)(1 + eps), F = A^2 / (q (1 - q)).
    Write q^_i = q_i + Eq, T^_i = T_i + Et (|Eq| <= q_i rq_i, |Et| <= T_i rT_i: the accumulated errors).  The
    chain between i and k adds only |k - i| more steps, so q^_k = q_k + Eq + lq, T^_k = T_k + Et + lt with
    |lq| <= q_k (D + 3) u, |lt| <= T_k (3 D + 5) u, D = |k - i|.  ln F_i - ln F_k therefore moves with (Eq, Et,
    the rounding of mu) only through the DIFFERENCE of the log-derivatives at the two bins."""
    with np.errstate(all="ignore"):
        d = mu1 - mu2
        A = q1 * q2 * d
        sig = q1 * q2 * d * d
        D = np.abs(ks - i).astype(np.float64)
        Ai, Ak = A[i], A[ks]
        dT = 2.0 * np.abs(1.0 / Ai - 1.0 / Ak)                                   # |d/dT (ln F_i - ln F_k)|
        dq = np.abs(-2.0 * M * (1.0 / Ai - 1.0 / Ak) - (1.0 / q1[i] - 1.0 / q1[ks]) + (1.0 / q2[i] - 1.0 / q2[ks]))
        dM = 2.0 * np.abs(q1[i] / Ai - q1[ks] / Ak)
        common = 1.5 * (dT * T[i] * rT[i] + dq * q1[i] * rq[i] + dM * M * 3.0 * U)
        lq = 0.0 if pow2 else q1[ks] * (D + 3.0) * U * INFL
        lt = T[ks] * (3.0 * D + 5.0) * U * INFL
        local = 1.5 * ((2.0 / np.abs(Ak)) * lt + (np.abs(2.0 * M / Ak) + 1.0 / q1[ks] + 1.0 / q2[ks]) * lq)

        def eps(j):     # roundings after (q^, T^): fl(q mu1), mu - ., 1 - q, the division, the difference, three products
            kappa = (np.abs(T[j] / (M - T[j])) + 4.0) * U
            return 7.0 * U + 2.0 * np.abs(mu2[j] / d[j]) * kappa

        def star_err(j):    # rounding of this model's own sigma (exact integers -> doubles)
            return 16.0 * U * (1.0 + (np.abs(mu1[j]) + np.abs(mu2[j])) / np.abs(d[j]))

        gap = (sig[i] - sig[ks]) / sig[i]
        need = (common + local + eps(i) + eps(ks) + star_err(i) + star_err(ks)) * INFL + 8.0 * U
        ok = (gap > need) & nonc[ks] & np.isfinite(need)
    return ok


def certify(h):
    h = np.asarray(h).astype(np.int64)
    n = len(h)
    N = int(h.sum())
    if N == 0:
        return True, 0, 0, -1
    idx = np.arange(n, dtype=np.int64)
    Cc = np.cumsum(h)
    S = np.cumsum(h * idx)
    ST = int(S[-1])
    nz = np.nonzero(h)[0]
    first, last = int(nz[0]), int(nz[-1])
    if ST >= 2 ** 53:
        return False, -1, -1, last
    pow2 = (N & (N - 1)) == 0
    Nf = float(N)
    slop = 4 * U
    invN = 1.0 / Nf          # x / N as x * (1 / N), like the kernel (exact for N = 2^k, else inside the slack)
    q1 = Cc.astype(np.float64) * invN
    q2 = (N - Cc).astype(np.float64) * invN
    rq = np.zeros(n) if pow2 else (np.maximum(idx - first + 1, 0).astype(np.float64) + 2.0) * U * INFL
    if pow2:   # the q1 chain is exact (multiples of 1/N), so are the skip decisions
        skip_front = q1 < EPS
        skip_tail = q1 > 1.0 - EPS
        nonc = ~(skip_front | skip_tail)
    else:
        skip_front = q1 * (1 + rq + slop) < EPS
        skip_tail = q1 * (1 - rq - slop) > 1.0 - EPS
        nonc = (q1 * (1 - rq - slop) >= EPS * (1 + slop)) & (q1 * (1 + rq + slop) <= (1.0 - EPS) * (1 - slop))
    live = idx >= first
    amb = ~(skip_front | skip_tail | nonc) & live
    nonc = nonc & live
    if not nonc.any():
        return (not amb.any()), 0, 0, last
    f = int(np.nonzero(nonc)[0][0])
    if amb[:f].any():
        return False, -1, -1, last
    S0 = int(S[f - 1]) if f > 0 else 0
    Sp_i = S - S0
    rest_i = ST - Sp_i
    rT = (3.0 * np.maximum(idx - f, 0).astype(np.float64) + 5.0) * U * INFL
    cand = (nonc | amb) & (idx >= f) & (idx <= last)
    with np.errstate(all="ignore"):
        mu_star = ST * invN
        B = Sp_i.astype(np.float64) * invN
        mu1 = Sp_i.astype(np.float64) / Cc.astype(np.float64)
        num = rest_i.astype(np.float64) * invN
        mu2 = rest_i.astype(np.float64) / (N - Cc).astype(np.float64)
        e_m1 = mu1 * ((rT + rq) * INFL + 2 * U)
        e_num = (mu_star * 3 * U + B * (rT + U) + U * (num + mu_star)) * INFL + 4 * U * mu_star
        e_q2 = (q1 * rq + U * q2) * INFL + 2 * U * q2
        den = q2 - e_q2
        e_mu2 = np.where(den > 0, (num + e_num) / den * (1 + 4 * U) - mu2 * (1 - 4 * U), np.inf)
        d = np.abs(mu1 - mu2)
        e_d = (e_m1 + e_mu2) * (1 + 4 * U) + 4 * U * (np.abs(mu1) + np.abs(mu2))
        up = q1 * (1 + rq) * (q2 + e_q2) * (d + e_d) * (d + e_d) * (1 + 16 * U)
        lo = q1 * (1 - rq) * np.maximum(q2 - e_q2, 0) * np.maximum(d - e_d, 0) * np.maximum(d - e_d, 0) * (1 - 16 * U)
    up = np.where(np.isnan(up), np.inf, up)
    up = np.where(cand, up, -1.0)
    lo = np.where(np.isnan(lo), 0.0, lo)
    lo = np.where(nonc & cand, lo, 0.0)
    istar = int(np.argmax(lo))
    Lmax = lo[istar]
    if not (Lmax > 0):
        return False, -1, -1, last
    comp = (up >= Lmax) & cand
    comp[istar] = False
    # second chance for the bins the interval test could not exclude: the accumulated errors of the two
    # chains at bin k and at istar are the SAME numbers up to the few roundings between the two bins
    # (common mode), and sigma reacts to them almost identically at neighbouring bins
    if comp.any():
        ks = np.nonzero(comp)[0]
        keep = ~differential_excludes(ks, istar, q1, q2, mu1, mu2, B, mu_star, rq, rT, nonc, pow2)
        comp[ks[~keep]] = False
    nc = int(comp.sum())
    kmax = max(istar, int(np.nonzero(comp)[0][-1])) if nc else istar
    return nc == 0, istar, nc, kmax
