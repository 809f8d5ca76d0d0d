"""A second, independent statement of the colour arithmetic in numpy float32 (every operation rounded separately by
construction: numpy has no fused multiply-add), against the C oracle -- guards the oracle against compiler contraction
or reordering (gcc -ffp-contract=off is a flag, this is a check).  Algorithm as restated from athas/matte (see
oracle/fs_oracle.h "parity unpinned") and fut/render_functions.fut:67-77, 95-105."""
import numpy as np

f32 = np.float32


def np_mix(m1, c1, m2, c2):
    m1, m2 = f32(m1), f32(m2)
    with np.errstate(all="ignore"):
        m12 = m1 + m2
        m1n, m2n = m1 / m12, m2 / m12
        out = 0
        for sh in (16, 8, 0):
            x1, x2 = f32((c1 >> sh) & 255) / f32(255), f32((c2 >> sh) & 255) / f32(255)
            v = np.sqrt(m1n * (x1 * x1) + m2n * (x2 * x2), dtype=f32)
            out |= np_channel(v) << sh
        a1, a2 = f32(c1 >> 24) / f32(255), f32(c2 >> 24) / f32(255)
        al = (m1 * a1 + m2 * a2) / m12
    return out | (np_channel(al) << 24)


def np_channel(x):
    x = f32(x)
    if x < 0:
        x = f32(0)
    elif x > 1:
        x = f32(1)
    v = x * f32(255)
    return 0 if np.isnan(v) else int(v)          # truncation; NaN converts to 0 (saturating conversion)


def test_mix_matches_numpy_float32(oracle):
    rng = np.random.default_rng(11)
    L = oracle.lib()
    for i in range(4000):
        c1, c2 = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
        kind = i % 4
        if kind == 0:                      # bilinear weights of a non-integer coordinate
            x = f32(rng.uniform(-3000, 3000))
            m1, m2 = np.ceil(x) - x, x - np.floor(x)
        elif kind == 1:                    # |coordinate| < 1: inexact weights
            x = f32(rng.uniform(-1, 1))
            m1, m2 = np.ceil(x) - x, x - np.floor(x)
        elif kind == 2:                    # the shadow bake's weights: 4 * count and 1
            m1, m2 = f32(4 * int(rng.integers(0, 255))), f32(1)
        else:                              # degenerate: both zero (integer coordinate), one zero
            m1, m2 = f32(0), f32(rng.choice([0.0, 1.0]))
        assert L.fso_mix(float(m1), c1, float(m2), c2) == np_mix(m1, c1, m2, c2), (float(m1), hex(c1), float(m2), hex(c2))


def test_bilinear_samplers_match_numpy_float32(oracle, c1w_d1):
    rgb, hgt = c1w_d1
    col = np.ascontiguousarray(rgb | np.uint32(0xFF000000))
    hgt = np.ascontiguousarray(hgt)
    q, r = hgt.shape
    rng = np.random.default_rng(12)
    L = oracle.lib()
    for _ in range(600):
        x, y = f32(rng.uniform(-2500, 2500)), f32(rng.uniform(-2500, 2500))
        fx, cx, fy, cy = np.floor(x), np.ceil(x), np.floor(y), np.ceil(y)
        x0, x1, y0, y1 = int(fx) % r, int(cx) % r, int(fy) % q, int(cy) % q     # Python % is the floored modulo
        wx0, wx1, wy0, wy1 = cx - x, x - fx, cy - y, y - fy
        xi1 = wx0 * f32(hgt[y0, x0]) + wx1 * f32(hgt[y0, x1])
        xi2 = wx0 * f32(hgt[y1, x0]) + wx1 * f32(hgt[y1, x1])
        want_h = wy0 * xi1 + wy1 * xi2
        got_h = L.fso_height_bilinear(hgt.ctypes.data, q, r, float(x), float(y), 0)
        assert f32(got_h) == want_h
        i1 = np_mix(wx0, int(col[y0, x0]), wx1, int(col[y0, x1]))
        i2 = np_mix(wx0, int(col[y1, x0]), wx1, int(col[y1, x1]))
        want_c = np_mix(wy0, i1, wy1, i2)
        assert L.fso_color_bilinear(col.ctypes.data, q, r, float(x), float(y), 0) == want_c
