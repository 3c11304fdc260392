"""The exact row-span computation of the rasterizer (ehb_row_span in easyhec_b200/csrc/ehb_kernels.cuh), restated
in numpy and checked against the brute-force test of every sample.

Per edge k the covered samples of a row are {dx : R_k + ax_k dx >= 0}.  With a = |ax|: ax < 0 bounds dx <= floor(R/a),
ax > 0 bounds dx >= -floor(R/a), ax == 0 is all or nothing.  floor(R/a) comes from ONE fp32 division estimate (the
device uses __fdividef, 2 ulp), clamped to +-(w+1), and ONE exact integer remainder; the claim is that this equals
testing every sample for all inputs the kernels produce (|quotient| <= w + 1 <= 8193, or beyond the row either way)."""
import numpy as np


def span_formula(R, ax, w, noise):
    lo = np.zeros(len(w), np.int64)
    hi = w - 1
    for k in range(3):
        a = np.abs(ax[k])
        a1 = np.where(a > 0, a, 1)
        q = (R[k].astype(np.float32) / a1.astype(np.float32)).astype(np.float32)
        q = (q * (1 + noise[k]).astype(np.float32)).astype(np.float32)       # the 2-ulp error of the fast division
        lim = (w + 1).astype(np.float32)
        q = np.minimum(np.maximum(q, -lim), lim)
        fl = np.floor(q).astype(np.int64)
        rem = R[k] - a1 * fl
        fl = fl + (rem >= a1) - (rem < 0)
        hik = np.where(ax[k] < 0, fl, np.where((ax[k] == 0) & (R[k] < 0), -1, w - 1))
        lok = np.where(ax[k] > 0, -fl, 0)
        hi = np.minimum(hi, hik)
        lo = np.maximum(lo, lok)
    return lo, hi


def span_brute(R, ax, w):
    lo = np.full(len(w), 10 ** 9)
    hi = np.full(len(w), -1)
    for dx in range(int(w.max())):
        ok = dx < w
        for k in range(3):
            ok &= (R[k] + ax[k] * dx >= 0)
        lo = np.where(ok & (lo == 10 ** 9), dx, lo)
        hi = np.where(ok, dx, hi)
    return lo, hi


def span_exact(R, ax, w):
    lo = np.zeros(len(w), np.int64)
    hi = w - 1
    for k in range(3):
        a = np.abs(ax[k])
        a1 = np.where(a > 0, a, 1)
        f = np.floor_divide(R[k], a1)
        hi = np.minimum(hi, np.where(ax[k] < 0, f, np.where((ax[k] == 0) & (R[k] < 0), -1, w - 1)))
        lo = np.maximum(lo, np.where(ax[k] > 0, -f, 0))
    return lo, hi


def _agree(a, b):
    (lo, hi), (blo, bhi) = a, b
    ea, eb = lo > hi, blo > bhi
    return not ((ea != eb) | (~eb & ((lo != blo) | (hi != bhi)))).any()


def test_small_rows_equal_brute_force():
    rng = np.random.RandomState(0)
    N = 100000
    w = rng.randint(1, 97, N).astype(np.int64)                                   # EHB_SMALL_AREA rows
    ax = [(-16 * rng.randint(-3000, 3001, N) * (rng.rand(N) > 0.05)).astype(np.int64) for _ in range(3)]
    R = [(rng.randint(-50, 150, N) * np.abs(ax[k]) // 8 + rng.randint(-40, 40, N) * (rng.rand(N) > 0.3)).astype(np.int64)
         for k in range(3)]
    for k in range(3):                                                           # exact multiples: samples on the edge
        m = rng.rand(N) < 0.2
        R[k] = np.where(m, np.abs(ax[k]) * rng.randint(-3, 100, N), R[k])
    noise = [rng.randint(-3, 4, N) * 2.0 ** -23 for _ in range(3)]
    lo, hi = span_formula(R, ax, w, noise)
    blo, bhi = span_brute(R, ax, w)
    assert ((bhi >= 0).sum()) > N // 3
    assert _agree((lo, hi), (np.where(bhi < 0, 1, blo), np.where(bhi < 0, 0, bhi)))


def test_large_values_and_clamped_quotients_equal_integer_floor_division():
    rng = np.random.RandomState(1)
    N = 50000
    w = rng.randint(1, 8161, N).astype(np.int64)                                 # up to the maximum resolution
    ax = [(-16 * rng.randint(-32767, 32768, N)).astype(np.int64) for _ in range(3)]   # |edge| < 2^15 sub-pixel units
    R = [rng.randint(-2 ** 30, 2 ** 30, N).astype(np.int64) for _ in range(3)]
    noise = [rng.randint(-3, 4, N) * 2.0 ** -23 for _ in range(3)]
    assert _agree(span_formula(R, ax, w, noise), span_exact(R, ax, w))
