"""The local occlusion bound of the column-parallel march (fsb_march_cols.cu, chunk_hidden) restated in numpy float32
and checked on the CPU: (1) the pyramid of local height maxima the library uploads (fsb_api.c build_height_pyramid)
against an independent numpy construction, (2) the bound itself -- for every sub-block of 4 depth steps x 8 columns the
row it yields must not lie below the row of any sample of the sub-block, where the samples are sampled and projected as
the reference does (fut/voxel_renderer.fut:63-66, :217-225, fut/render_functions.fut:67-77) in float32.  No GPU needed."""
import ctypes
import math

import numpy as np
import pytest

f32 = np.float32


def np_pyramid(h):
    """level L -> array (q >> L, r >> L): highest texel of the 2 x 2 blocks of 2^L texels starting at block (by, bx), wrapped"""
    levels = {}
    cur = h.astype(np.uint8)
    L = 0
    while cur.shape[0] >= 2 and cur.shape[1] >= 2:
        cur = np.maximum(np.maximum(cur[0::2, 0::2], cur[1::2, 0::2]), np.maximum(cur[0::2, 1::2], cur[1::2, 1::2]))
        L += 1
        d = np.maximum(np.maximum(cur, np.roll(cur, -1, 0)), np.maximum(np.roll(cur, -1, 1), np.roll(np.roll(cur, -1, 0), -1, 1)))
        levels[L] = d
    return levels


@pytest.mark.parametrize("shape", [(2, 2), (2, 8), (4, 4), (64, 256), (512, 128), (1024, 1024)])
def test_pyramid_builder_matches_numpy(fsb, shape):
    q, r = shape
    rng = np.random.default_rng(q * 131 + r)
    h = rng.integers(0, 256, size=(q, r)).astype(np.int32)
    n = q * r
    out = np.zeros(n // 3 + 16, np.uint8)
    f = fsb.lib().fsb_debug_height_pyramid
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    levels = f(h.ctypes.data, q, r, out.ctypes.data, out.size)
    want = np_pyramid(h)
    assert levels == len(want) == int(math.log2(min(q, r)))
    for L, d in want.items():
        off = (n - (n >> (2 * L - 2))) // 3
        got = out[off:off + d.size].reshape(d.shape)
        assert np.array_equal(got, d), L
        # what the level promises: entry (by, bx) bounds every texel of [by 2^L, (by + 2) 2^L) x [bx 2^L, (bx + 2) 2^L)
        by, bx = int(rng.integers(0, d.shape[0])), int(rng.integers(0, d.shape[1]))
        ys = np.arange(by << L, (by + 2) << L) % q
        xs = np.arange(bx << L, (bx + 2) << L) % r
        assert got[by, bx] == h[np.ix_(ys, xs)].max()
    assert f(h.ctypes.data, 3, 5, None, 0) == -1          # power-of-two maps only


def depth_table(cam, prm, w, n_z):
    """fsb_device.cuh depth_entry / fut/voxel_renderer.fut:28-34,43-60,217 in float32"""
    s, c, view = f32(np.sin(f32(cam[3]))), f32(np.cos(f32(cam[3]))), f32(cam[6])
    sv, cv = s * view, c * view
    a_lx, a_ly, a_rx, a_ry = -c - sv, s - cv, c - sv, -s - cv
    i = np.arange(1, n_z + 1, dtype=f32)
    with np.errstate(all="ignore"):
        z = (i / f32(2)) * (f32(2) * f32(prm["z0"]) + (i - f32(1)) * f32(prm["delta"]))
        lx, ly, rx, ry = a_lx * z, a_ly * z, a_rx * z, a_ry * z
        dx, dy = (rx - lx) / f32(w), (ry - ly) / f32(w)
        sx, sy = lx + f32(cam[0]), ly + f32(cam[1])
        iz = (f32(1) / z) * f32(w // 2)
    return sx, sy, dx, dy, iz


def sample_rows(hgt, sx, sy, dx, dy, iz, cam_h, horizon, cols):
    q, r = hgt.shape
    j = cols.astype(f32)[None, :]
    with np.errstate(all="ignore"):
        x = sx[:, None] + j * dx[:, None]
        y = sy[:, None] + j * dy[:, None]
        fx, cx, fy, cy = np.floor(x), np.ceil(x), np.floor(y), np.ceil(y)
        x0, x1 = fx.astype(np.int64) % r, cx.astype(np.int64) % r
        y0, y1 = fy.astype(np.int64) % q, cy.astype(np.int64) % q
        wx0, wx1, wy0, wy1 = cx - x, x - fx, cy - y, y - fy
        H = hgt.astype(f32)
        xi1 = wx0 * H[y0, x0] + wx1 * H[y0, x1]
        xi2 = wx0 * H[y1, x0] + wx1 * H[y1, x1]
        height = wy0 * xi1 + wy1 * xi2
        rel = (f32(cam_h) - height) * iz[:, None] + f32(horizon)
        rows = np.where(np.isnan(rel), 0, np.clip(np.trunc(np.clip(rel, -3e9, 3e9)), 0, 2 ** 31 - 1)).astype(np.int64)
    return x, y, rows


def round_up(v64):
    v = f32(v64)
    return np.where(v.astype(np.float64) < v64, np.nextafter(v, f32(np.inf)), v)


def round_down(v64):
    v = f32(v64)
    return np.where(v.astype(np.float64) > v64, np.nextafter(v, f32(-np.inf)), v)


@pytest.mark.parametrize("terrain", ["fbm", "spikes", "steep"])
def test_local_bound_never_exceeds_a_sample_row(fsb, terrain):
    rng = np.random.default_rng(5)
    if terrain == "fbm":
        _, hgt = fsb.terrain_fbm(512)
        hgt = (hgt & 0xFF).astype(np.int32)
    elif terrain == "spikes":
        hgt = np.zeros((256, 512), np.int32)
        hgt[rng.integers(0, 256, 400), rng.integers(0, 512, 400)] = 255
    else:
        yy, xx = np.mgrid[0:128, 0:128]
        hgt = (128 + 120 * np.sin(xx / 5.0) * np.cos(yy / 3.0)).astype(np.int32)
    q, r = hgt.shape
    pyr = np_pyramid(hgt)
    levels = len(pyr)
    w = 256
    prm = dict(z0=0.0, delta=0.001)
    tight = total = 0
    for cam in [(256.37, 130.73, 90, 2.2, 120, 700, 1.2), (100.0, 7.0, 30, -0.7, 150, 600, 0.8), (3.25, 90.5, 300, 4.0, -20, 900, 1.5),
                (-2000.5, 77777.25, 200, 9.0, 200, 500, 2.0), (40.5, 40.5, 64, 1.0, 200, 300, 1.2), (12.3, 45.6, 5, 0.1, 100, 1200, 0.6)]:
        n_z = len(fsb.get_zs(prm["delta"], cam[5], prm["z0"]))
        n_z -= n_z % 32
        sx, sy, dx, dy, iz = depth_table(cam, prm, w, n_z)
        cam_h, horizon = cam[2], cam[4]
        x, y, rows = sample_rows(hgt, sx, sy, dx, dy, iz, cam_h, horizon, np.arange(w))
        for g in range(w // 32):
            for k0 in range(0, n_z, 4):
                for b in range(4):
                    ja, jb = f32(g * 32 + b * 8), f32(g * 32 + b * 8 + 7)
                    k1 = k0 + 3
                    with np.errstate(all="ignore"):
                        cxs = [sx[k0] + ja * dx[k0], sx[k0] + jb * dx[k0], sx[k1] + ja * dx[k1], sx[k1] + jb * dx[k1]]
                        cys = [sy[k0] + ja * dy[k0], sy[k0] + jb * dy[k0], sy[k1] + ja * dy[k1], sy[k1] + jb * dy[k1]]
                    mag = sum(abs(float(v)) for v in cxs + cys)
                    if not mag < 3.2e7:
                        continue
                    x0, x1 = math.floor(min(cxs)) - 2, math.floor(max(cxs)) + 3
                    y0, y1 = math.floor(min(cys)) - 2, math.floor(max(cys)) + 3
                    ext = max(x1 - x0, y1 - y0) + 1
                    L = (ext - 1).bit_length()
                    if L > levels:
                        continue
                    d_ = pyr[L]
                    hm = d_[(y0 >> L) % d_.shape[0], (x0 >> L) % d_.shape[1]]
                    hb = round_up(float(hm) + float(f32(0.501)))
                    dd = round_down(float(f32(cam_h)) - float(hb))
                    izs = iz[k1] if dd >= 0 else iz[k0]
                    with np.errstate(all="ignore"):
                        rel = f32(dd) * izs + f32(horizon)
                    bound = 0 if np.isnan(rel) else int(max(0, min(float(np.trunc(rel)), 2 ** 31 - 1)))
                    block = rows[k0:k0 + 4, g * 32 + b * 8:g * 32 + b * 8 + 8]
                    # every texel the sub-block's samples read lies inside the box the pyramid entry covers
                    xs, ys = x[k0:k0 + 4, g * 32 + b * 8:g * 32 + b * 8 + 8], y[k0:k0 + 4, g * 32 + b * 8:g * 32 + b * 8 + 8]
                    assert np.floor(xs).min() >= x0 and np.ceil(xs).max() <= x1 and np.floor(ys).min() >= y0 and np.ceil(ys).max() <= y1
                    assert bound <= block.min(), (cam, g, k0, b, bound, int(block.min()), int(hm), L)
                    total += 1
                    tight += bound > 0
    assert total > 1000 and tight > 0
