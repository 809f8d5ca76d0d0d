"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same
inputs.  Bar: bit-exact (u32 frames; integer/byte work).  Run on the GPU box with `-m gpu`."""
import math

import numpy as np
import pytest

from conftest import SKY

pytestmark = pytest.mark.gpu


def ocam(oracle, cam):
    return oracle.Camera(cam.x, cam.y, cam.height, cam.angle, cam.horizon, cam.distance, cam.fov, cam.sky_color)


def oprm(oracle, p):
    return oracle.Params(p.z0, p.delta, p.invz_param1, p.invz_param2, p.filter, p.sentinel, p.f2i_mode,
                         1 if p.flags & 8 else 0)  # FSB_FLAG_SMOOTHING -> fso_params.smoothing


def check(fsb, oracle, ctx, mp, color, height, cam, prm, h, w, masked=True):
    got = ctx.render(cam, prm, mp, h, w)
    hm = height & 0xFF if masked else height
    want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), color, hm, h, w)
    nbad = int((got != want).sum())
    assert nbad == 0, "%d / %d pixels differ" % (nbad, h * w)
    return got


def test_native_library_loaded(fsb, gpu_ctx):
    assert "B200" in gpu_ctx.device_name, gpu_ctx.device_name
    maps = open("/proc/self/maps").read()
    assert "libfutspace_b200.so" in maps


def test_branch_free_sqrt_is_correctly_rounded(gpu_ctx):
    # every float in [2^-100, 2.0) (bit patterns 0x0D800000 .. 0x40000000) and zero: identical to sqrt.rn.f32
    assert gpu_ctx.selftest_sqrt(0x0D800000, 0x40000000) == 0
    assert gpu_ctx.selftest_sqrt(0, 1) == 0


def test_tests_variant_golden(fsb, oracle, gpu_ctx, c1w_d1, golden_frames):
    # tests/futspace.fut main: C1W/D1, fixed camera, 400x800 (colours without alpha -> packed path with alpha 0)
    rgb, hgt = c1w_d1
    mp = gpu_ctx.upload_map(rgb, hgt)
    assert mp.packed
    cam = fsb.Camera(512, 800, 78, 0, 100, 800, 1, SKY)
    n0 = gpu_ctx.launch_count
    got = check(fsb, oracle, gpu_ctx, mp, rgb, hgt, cam, fsb.tests_variant_params(), 400, 800)
    assert gpu_ctx.launch_count - n0 == 3   # set-up, march, expand
    assert np.array_equal(got, golden_frames["tests_variant_400x800"])
    mp.free()


def test_live_variant_golden_config1(fsb, oracle, gpu_ctx, c1w_d1, golden_frames):
    # BASELINE config 1: 1024x1024 converted maps, 1024x768 frame, distance 1000, init camera
    rgb, hgt = c1w_d1
    col = rgb | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt)
    cam = fsb.Camera(0.98, 0.6, 58, 2.2, 200, 1000, 1.2, SKY)
    got = check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(), 768, 1024)
    assert np.array_equal(got, golden_frames["live_init_768x1024_d1000"])
    mp.free()


POSES = [
    (512.37, 512.73, 180, 2.2, 230, 1000, 1.2),   # bench pose shape
    (100.0, 7.0, 30, -0.7, 300, 600, 0.8),        # integer coordinates (fact 9), camera below terrain
    (3.25, 900.5, 260, 4.0, -20, 700, 1.5),       # horizon < 0
    (-0.4, 0.3, 120, 0.3, 900, 500, 1.2),         # |coordinate| < 1: inexact bilinear weights; horizon > h
    (-2000.5, 77777.25, 200, 9.0, 384, 800, 2.0), # far outside the map: floored-modulo wrap
    (512.5, 512.5, 64, 1.0, 400, 400, 1.2),       # camera height == water level: NaN/inf at z = 0
    (5.0e6, -7.25e6, 200, 0.5, 384, 300, 1.2),    # beyond the fast paths' coordinate range: generic kernel takes over
    (4194000.5, 100.25, 200, 2.0, 384, 300, 1.2), # just inside / across the range bound
]


@pytest.mark.parametrize("filt", [1, 0])
@pytest.mark.parametrize("sentinel", [0, 1])
@pytest.mark.parametrize("flags", [0, 2, 4, 16, 16 | 4], ids=["texture", "tiled_ldg", "texture_nocull", "texture_march_z", "march_z_nocull"])
def test_fbm_poses_packed(fsb, oracle, gpu_ctx, fbm1024, filt, sentinel, flags):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    assert mp.packed
    prm = fsb.default_params(filter=filt, sentinel=sentinel, flags=flags)
    for p in POSES:
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 768, 1024)
    mp.free()


@pytest.mark.parametrize("filt", [1, 0])
@pytest.mark.parametrize("f2i", [0, 1, 2])
def test_generic_kernel_all_f2i_modes(fsb, oracle, gpu_ctx, fbm1024, filt, f2i):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params(filter=filt, f2i_mode=f2i, flags=fsb.FLAG_FORCE_GENERIC)
    for p in POSES[:3]:
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 300, 400)
    # z0 > 0: all three conversions agree and produce terrain
    prm = fsb.tests_variant_params(filter=filt, f2i_mode=f2i, flags=fsb.FLAG_FORCE_GENERIC)
    out = check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*POSES[0], SKY), prm, 300, 400)
    assert len(np.unique(out)) > 50
    mp.free()


def test_unpackable_maps(fsb, oracle, gpu_ctx):
    # non-power-of-two, heights > 255 unmasked, varying alpha -> two-plane generic kernel
    rng = np.random.default_rng(3)
    q, r = 300, 517
    yy, xx = np.mgrid[0:q, 0:r]
    hgt = (200 + 180 * np.sin(xx / 37.0) * np.cos(yy / 23.0)).astype(np.int32)
    col = rng.integers(0, 1 << 32, size=(q, r), dtype=np.uint64).astype(np.uint32)
    mp = gpu_ctx.upload_map(col, hgt, mask_heights=False)
    assert not mp.packed
    for filt in (0, 1):
        prm = fsb.default_params(filter=filt)
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(150.3, 100.6, 420, 0.9, 150, 500, 1.2, SKY), prm, 240, 333,
              masked=False)
    mp.free()
    # same map, masked: update_map semantics (fut/interactive.fut:189)
    mp = gpu_ctx.upload_map(col, hgt, mask_heights=True)
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(150.3, 100.6, 300, 0.9, 150, 500, 1.2, SKY),
          fsb.default_params(), 240, 333, masked=True)
    mp.free()


@pytest.mark.parametrize("filt", [1, 0])
def test_packed_texture_path_any_map_size(fsb, oracle, gpu_ctx, filt):
    # packable content (heights <= 255, uniform alpha) on a non-power-of-two map: texture path with hardware wrap
    rng = np.random.default_rng(11)
    q, r = 300, 517
    yy, xx = np.mgrid[0:q, 0:r]
    hgt = (120 + 100 * np.sin(xx / 37.0) * np.cos(yy / 23.0)).astype(np.int32)
    col = (rng.integers(0, 1 << 24, size=(q, r), dtype=np.uint64).astype(np.uint32)) | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt)
    assert mp.packed
    prm = fsb.default_params(filter=filt)
    for p in ((150.3, 100.6, 260, 0.9, 150, 500, 1.2), (-777.25, 5000.5, 230, 3.3, 100, 400, 1.0), (516.5, 299.5, 250, 5.0, 120, 300, 1.4),
              (999000.25, -998500.5, 240, 1.1, 120, 300, 1.2),    # largest coordinates the texture path takes on a non-power-of-two map
              (1.5e6, 2.5e6, 240, 1.1, 120, 300, 1.2)):           # beyond: generic kernel
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 240, 333)
    mp.free()


def test_ragged_and_degenerate_sizes(fsb, oracle, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params()
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 40, 300, 1.2, SKY)
    for h, w in ((1, 1), (7, 5), (33, 9), (100, 17), (31, 64), (257, 8)):
        check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, h, w)
    # n_z = 0 -> all sky ; n_z = 1 ; n_z = 33 (one full chunk + 1)
    for dist in (0.0004, 0.001, 0.6):
        cam.distance = dist
        out = check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, 64, 48)
    cam.distance = 0.0004
    assert (gpu_ctx.render(cam, prm, mp, 16, 16) == SKY).all()
    mp.free()


def test_error_behaviour(fsb, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    cam = fsb.Camera(1, 2, 100, 0, 50, -10.0, 1.2, SKY)   # sqrt of a negative number: z-series undefined
    with pytest.raises(fsb.FsbError) as e:
        gpu_ctx.render(cam, fsb.default_params(), mp, 64, 64)
    assert e.value.code == fsb.ERR_RANGE and "z-series" in str(e.value)
    cam.distance = 100
    with pytest.raises(fsb.FsbError) as e:
        gpu_ctx.render(cam, fsb.default_params(filter=7), mp, 64, 64)
    assert e.value.code == fsb.ERR_ARG
    with pytest.raises(fsb.FsbError) as e:
        gpu_ctx.render(cam, fsb.default_params(), mp, 40000, 64)   # taller than the supported maximum
    assert e.value.code == fsb.ERR_RANGE
    # the context stays usable after an error
    gpu_ctx.render(cam, fsb.default_params(), mp, 64, 64)
    # a bad pose late in a host batch is reported before anything is queued: the output buffer is left untouched
    cams = [fsb.Camera(1, 2, 100, 0, 50, 100.0, 1.2, SKY) for _ in range(40)]
    cams[37].distance = -5.0
    out = np.full((40, 1024, 1024), 0xDEADBEEF, np.uint32)      # 4 MiB frames: the batch would span three staging chunks
    with pytest.raises(fsb.FsbError) as e:
        gpu_ctx.render_batch(cams, fsb.default_params(), mp, 1024, 1024, out=out)
    assert e.value.code == fsb.ERR_RANGE and "pose 37" in str(e.value)
    assert (out == 0xDEADBEEF).all()
    mp.free()


def camera_path(fsb, m, n, dist):
    cams = []
    for i in range(n):
        th = 2 * math.pi * i / n
        cams.append(fsb.Camera(m / 2 + m / 4 * math.cos(th), m / 2 + m / 4 * math.sin(th),
                               160 + 40 * math.sin(2 * th), 2.2 + th, 300, dist, 1.2, SKY))
    return cams


def test_batch_equals_single_frames(fsb, oracle, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params()
    cams = camera_path(fsb, 1024, 12, 700)
    cams[3].distance = 300   # ragged z-series inside one batch
    batch = gpu_ctx.render_batch(cams, prm, mp, 270, 480)
    for i, cam in enumerate(cams):
        single = gpu_ctx.render(cam, prm, mp, 270, 480)
        assert np.array_equal(batch[i], single)
        if i % 4 == 0:
            want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt, 270, 480)
            assert np.array_equal(single, want)
    mp.free()


def test_device_output_and_column_split(fsb, oracle, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params()
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 200, 900, 1.2, SKY)
    h, w = 540, 1000
    want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt, h, w)
    dev = gpu_ctx.device_malloc(h * w * 4)
    gpu_ctx.render_device(cam, prm, mp, h, w, dev)
    assert np.array_equal(gpu_ctx.download(dev, (h, w)), want)
    # column slabs written straight into one row-major frame (the column-split multi-GPU layout)
    dev2 = gpu_ctx.device_malloc(h * w * 4)
    bounds = [0, 123, 500, 504, 1000]
    for a, b in zip(bounds[:-1], bounds[1:]):
        gpu_ctx.render_columns_device(cam, prm, mp, h, w, a, b, dev2 + 4 * a, w)
    assert np.array_equal(gpu_ctx.download(dev2, (h, w)), want)
    # a slab on its own (gather layout [h][ncols])
    slab = gpu_ctx.device_malloc(h * 377 * 4)
    gpu_ctx.render_columns_device(cam, prm, mp, h, w, 123, 500, slab, 0)
    assert np.array_equal(gpu_ctx.download(slab, (h, 377)), want[:, 123:500])
    for p in (dev, dev2, slab):
        gpu_ctx.device_free(p)
    mp.free()


def test_full_size_4k_distance_4000(fsb, oracle, gpu_ctx):
    # BASELINE config 3 at full size: 3840x2160, 4096^2 fBm, distance 4000 -- bit-exact against the oracle
    col, hgt = fsb.terrain_fbm(4096)
    mp = gpu_ctx.upload_map(col, hgt)
    m = 4096
    cam = fsb.Camera(m / 2 + 0.37, m / 2 + 0.73, 200, 2.2, 0.3 * 2160, 4000, 1.2, SKY)
    got = check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(), 2160, 3840)
    # size-independent properties: rendering is idempotent, and every column is "first hit wins":
    again = gpu_ctx.render(cam, fsb.default_params(), mp, 2160, 3840)
    assert np.array_equal(got, again)
    assert (got != SKY).mean() > 0.3
    mp.free()


def test_full_size_1080p_distance_2000(fsb, oracle, gpu_ctx):
    col, hgt = fsb.terrain_fbm(2048)
    mp = gpu_ctx.upload_map(col, hgt)
    m = 2048
    cam = fsb.Camera(m / 2 + 0.37, m / 2 + 0.73, 200, 2.2, 0.3 * 1080, 2000, 1.2, SKY)
    check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(), 1080, 1920)
    mp.free()


@pytest.mark.parametrize("sun_height,sun_ang", [(0.1, 0.1), (1.3, 0.4), (1.45, -2.0)])
def test_shadow_bake_matches_oracle(fsb, oracle, gpu_ctx, c1w_d1, sun_height, sun_ang):
    # update_map's shadow bake (fut/interactive.fut:194-196, fut/effects.fut:108-125) on the reference's own map pair
    rgb, hgt = c1w_d1
    col = rgb | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt)
    sun = fsb.sun_vector(sun_height, sun_ang)
    got = gpu_ctx.bake_shadows(mp, sun, 1024, 1024)
    want = oracle.bake_shadows(col, hgt, sun, 1024, 1024)
    assert np.array_equal(got, want)
    if sun_height > 1.0:
        assert (got != col).mean() > 0.05      # a low sun really casts shadows
    mp.free()
    # the baked map is what render draws (lsc.shadowed_color, fut/interactive.fut:181)
    mp2 = gpu_ctx.upload_map(got, hgt)
    cam = fsb.Camera(0.98, 0.6, 58, 2.2, 200, 800, 1.2, SKY)
    check(fsb, oracle, gpu_ctx, mp2, got, hgt, cam, fsb.default_params(), 384, 512)
    mp2.free()


def test_shadow_bake_non_square_output(fsb, oracle, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col[:300, :517].copy(), hgt[:300, :517].copy())
    sun = fsb.sun_vector(1.2, 0.9)
    got = gpu_ctx.bake_shadows(mp, sun, 200, 333)
    want = oracle.bake_shadows(col[:300, :517].copy(), hgt[:300, :517].copy(), sun, 200, 333)
    assert np.array_equal(got, want)
    mp.free()


def test_random_poses_fuzz(fsb, oracle, gpu_ctx, fbm1024, c1w_d1):
    """Seeded fuzz over cameras, renderer constants, frame sizes and kernel variants; bit-exact or fail."""
    rng = np.random.default_rng(20261017)
    col, hgt = fbm1024
    rgb, h2 = c1w_d1
    maps = [(col, hgt, gpu_ctx.upload_map(col, hgt)), (rgb | 0xFF000000, h2, gpu_ctx.upload_map(rgb | 0xFF000000, h2))]
    for it in range(48):
        cmap, hmap, mp = maps[it % 2]
        h, w = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        cam = fsb.Camera(float(rng.uniform(-3000, 3000)), float(rng.uniform(-3000, 3000)), float(rng.uniform(0, 400)),
                         float(rng.uniform(-7, 7)), float(rng.uniform(-100, h + 100)), float(rng.uniform(0.5, 1500)),
                         float(rng.uniform(0.3, 2.5)), int(rng.integers(0, 1 << 32)))
        if it % 7 == 0:   # exact integers / half-integers provoke the degenerate weights (fact 9)
            cam.x, cam.y, cam.angle = float(int(cam.x)), float(int(cam.y)) + 0.5, 0.0
        prm = fsb.default_params() if it % 3 else fsb.tests_variant_params()
        prm.filter = int(rng.integers(0, 2))
        prm.sentinel = int(rng.integers(0, 2))
        prm.flags = int(rng.choice([0, 0, 0, 1, 2, 4]))
        if it % 5 == 0:
            prm.invz_param1, prm.invz_param2 = float(rng.uniform(0.2, 3.0)), float(rng.uniform(20, 600))
        if it % 11 == 0:
            prm.z0, prm.delta = float(rng.uniform(0.0, 3.0)), float(rng.uniform(0.0005, 0.02))
        check(fsb, oracle, gpu_ctx, mp, cmap, hmap, cam, prm, h, w)
    for _, _, mp in maps:
        mp.free()


def test_effects_post_passes(fsb, oracle, gpu_ctx, fbm1024):
    # fut/effects.fut:27-52 on a rendered frame (the reference never calls them; restated and matched anyway)
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 90, 600, 1.2, SKY)
    frame = gpu_ctx.render(cam, fsb.default_params(), mp, 200, 333)
    assert np.array_equal(gpu_ctx.effect_interpolate(frame), oracle.interpolate2(frame))
    for pd in (1, 3):
        assert np.array_equal(gpu_ctx.effect_interpolate(frame, pd), oracle.interpolate(pd, frame))
    with pytest.raises(fsb.FsbError) as e:      # h > w: the reference's `% h` on x would read out of bounds
        gpu_ctx.effect_interpolate(frame[:, :100].copy(), 1)
    assert e.value.code == fsb.ERR_RANGE
    mp.free()


def test_full_size_8k_in_column_slabs(fsb, oracle, gpu_ctx):
    # BASELINE config 5 frame (7680x4320, distance 4000) assembled from 4 column slabs written into one frame;
    # the map is 4096^2 here (the 16384^2 one is exercised by bench.py --workload 8k-colsplit)
    from futspace_b200.shard import column_bounds
    col, hgt = fsb.terrain_fbm(4096)
    mp = gpu_ctx.upload_map(col, hgt)
    h, w = 4320, 7680
    cam = fsb.Camera(2048.37, 2048.73, 210, 2.2, 0.3 * h, 4000, 1.2, SKY)
    prm = fsb.default_params()
    dev = gpu_ctx.device_malloc(h * w * 4)
    b = column_bounds(w, 4)
    for r in range(4):
        gpu_ctx.render_columns_device(cam, prm, mp, h, w, b[r], b[r + 1], dev + 4 * b[r], w)
    got = gpu_ctx.download(dev, (h, w))
    want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt, h, w)
    assert np.array_equal(got, want)
    gpu_ctx.device_free(dev)
    mp.free()


@pytest.mark.parametrize("shape", [(1, 1), (2, 1), (3, 5), (4, 8), (7, 64)])
def test_tiny_maps(fsb, oracle, gpu_ctx, shape):
    # down to the reference's dummy landscape [[0],[0]] (fut/interactive.fut:38-43): everything wraps
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    hgt = rng.integers(0, 256, size=shape).astype(np.int32)
    col = rng.integers(0, 1 << 24, size=shape, dtype=np.uint64).astype(np.uint32) | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt)
    for filt in (0, 1):
        for cam in (fsb.Camera(0.98, 0.6, 300, 2.2, 40, 200, 1.2, SKY), fsb.Camera(-3.5, 2.25, 120, 0.4, 60, 150, 0.9, SKY)):
            check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(filter=filt), 96, 128)
    mp.free()


def test_huge_unmasked_heights(fsb, oracle, gpu_ctx):
    # i32 heights around +-1e9 (two-plane generic kernel, occlusion bound active): f32 rounding of the interpolation is
    # tens of units here, the bound's margin must cover it
    rng = np.random.default_rng(5)
    q, r = 256, 256
    yy, xx = np.mgrid[0:q, 0:r]
    hgt = (1.0e9 * np.sin(xx / 17.0) * np.cos(yy / 29.0)).astype(np.int32)
    col = rng.integers(0, 1 << 24, size=(q, r), dtype=np.uint64).astype(np.uint32) | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt, mask_heights=False)
    assert not mp.packed
    for cam_h in (9.99e8, 1.2e9, -3.0e8):
        for filt in (0, 1):
            prm = fsb.default_params(filter=filt, invz_param1=2.0e-7)
            cam = fsb.Camera(100.3, 77.7, cam_h, 0.9, 120, 400, 1.2, SKY)
            a = check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, 240, 200, masked=False)
            prm.flags = fsb.FLAG_NO_CULL
            b = gpu_ctx.render(cam, prm, mp, 240, 200)
            assert np.array_equal(a, b)
    mp.free()


@pytest.mark.parametrize("filt", [1, 0])
@pytest.mark.parametrize("flags", [8, 8 | 1, 8 | 2, 8 | 4, 8 | 16], ids=["texture", "generic", "tiled_ldg", "texture_nocull", "texture_march_z"])
def test_smoothing_on(fsb, oracle, gpu_ctx, fbm1024, c1w_d1, filt, flags):
    """Smoothing #on (fut/voxel_renderer.fut:175-213) under the sequential-scatter semantics the oracle states."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params(filter=filt, flags=flags)
    for p in POSES[:6]:
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 300, 400)
    # ragged sizes: partial bands, a single column, a single row
    for h, w in ((1, 1), (33, 5), (95, 64), (257, 31)):
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*POSES[0][:4], 0.3 * h, 500, 1.2, SKY), prm, h, w)
    mp.free()
    rgb, h2 = c1w_d1
    mp = gpu_ctx.upload_map(rgb | 0xFF000000, h2)
    cam = fsb.Camera(0.98, 0.6, 58, 2.2, 200, 1000, 1.2, SKY)   # the demo's initial pose, fut/interactive.fut:28-36
    on = check(fsb, oracle, gpu_ctx, mp, rgb | 0xFF000000, h2, cam, prm, 768, 1024)
    off = gpu_ctx.render(cam, fsb.default_params(filter=filt), mp, 768, 1024)
    assert 0 < (on != off).mean() < 0.5          # the mode changes the picture, and only the blended spans
    # tests-variant constants (z0 = 1: no z = 0 trap) still with the zero sentinel
    prm2 = fsb.tests_variant_params(filter=filt, sentinel=0, flags=flags)
    check(fsb, oracle, gpu_ctx, mp, rgb | 0xFF000000, h2, fsb.Camera(512, 800, 78, 0, 100, 800, 1, SKY), prm2, 400, 800)
    mp.free()


def test_smoothing_batch_and_errors(fsb, oracle, gpu_ctx, fbm1024):
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params(flags=fsb.FLAG_SMOOTHING)
    cams = [fsb.Camera(*p, SKY) for p in POSES[:5]]
    frames = gpu_ctx.render_batch(cams, prm, mp, 200, 320)
    for cam, got in zip(cams, frames):
        want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt & 0xFF, 200, 320)
        assert np.array_equal(got, want)
    with pytest.raises(fsb.FsbError):   # smoothing exists only in fut/voxel_renderer.fut (zero sentinel)
        gpu_ctx.render(cams[0], fsb.default_params(sentinel=1, flags=fsb.FLAG_SMOOTHING), mp, 64, 64)
    mp.free()


@pytest.mark.parametrize("filt", [1, 0])
def test_occlusion_bound_needs_positive_decreasing_inv_z(fsb, oracle, gpu_ctx, fbm1024, filt):
    """The terrain-height bound and the max-tap pre-test assume inv_z > 0 and decreasing along the depth series.
    A negative invz_param1 flips the projection, a negative z0 makes the first depths negative
    (fut/voxel_renderer.fut:33, :217): both must switch the shortcuts off, not produce different frames."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    for kw in (dict(invz_param1=-1.0), dict(invz_param1=-0.3, invz_param2=200.0), dict(z0=-0.4), dict(z0=-2.0, delta=0.01),
               dict(invz_param1=0.0)):
        prm = fsb.default_params(filter=filt, **kw)
        for p in POSES[:4]:
            check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 240, 320)
    mp.free()


def test_no_device_memory_leak_across_maps_contexts_and_sizes(fsb, gpu_ctx, fbm1024):
    """Maps, contexts and the per-context scratch (depth tables, record lists, staging frames) are released: device
    memory in use returns to where it started after many create / render / free cycles at changing sizes."""
    import torch
    col, hgt = fbm1024
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 90, 400, 1.2, SKY)

    def cycle(n):
        for i in range(n):
            ctx = fsb.Context(0)
            mp = ctx.upload_map(col[: 256 + 64 * (i % 3), : 512 - 32 * (i % 2)].copy(), hgt[: 256 + 64 * (i % 3), : 512 - 32 * (i % 2)].copy())
            h, w = 120 + 40 * (i % 4), 200 + 56 * (i % 3)
            ctx.render(cam, fsb.default_params(flags=(i % 2) * fsb.FLAG_SMOOTHING), mp, h, w)
            ctx.render_batch([cam] * 3, fsb.default_params(filter=i % 2), mp, h, w)
            mp.free()
            ctx.close()

    cycle(3)                                   # warm the allocator / module state
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    cycle(24)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), "device memory in use grew by %.1f MiB" % ((free0 - free1) / 2**20)
    # same within one context: growing and shrinking frames must reuse or release the scratch, maps must be freed
    before, _ = torch.cuda.mem_get_info()
    for i in range(20):
        mp = gpu_ctx.upload_map(col, hgt)
        gpu_ctx.render(cam, fsb.default_params(), mp, 64 + 16 * (i % 5), 96)
        mp.free()
    after, _ = torch.cuda.mem_get_info()
    assert before - after < (64 << 20)         # the context keeps its (bounded) scratch, nothing per map


def test_long_draw_distances(fsb, oracle, gpu_ctx, fbm1024):
    """Depth series far longer than the bench configs (n_z = 10 954 and 31 622): table sizing, culling prefix search over
    many chunks, record indices; plus the sample-index limit of the smoothing record word."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    for dist, h, w in ((60000.0, 96, 64), (500000.0, 40, 33)):
        for flags in (0, fsb.FLAG_NO_CULL, fsb.FLAG_SMOOTHING):
            cam = fsb.Camera(512.37, 512.73, 400 if dist > 1e5 else 120, 2.2, 0.4 * h, dist, 1.2, SKY)
            check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(flags=flags), h, w)
    assert len(fsb.get_zs(0.001, 1.0e7, 0.0)) > (1 << 17)
    with pytest.raises(fsb.FsbError):      # 141 421 samples do not fit the 17-bit sample index of a smoothing record
        gpu_ctx.render(fsb.Camera(512.37, 512.73, 120, 2.2, 20, 1.0e7, 1.2, SKY), fsb.default_params(flags=fsb.FLAG_SMOOTHING),
                       mp, 32, 16)
    mp.free()


def test_real_map_pairs(fsb, oracle, gpu_ctx, real_maps):
    """the reference's own map pairs at the tests/futspace.fut camera (tests variant: nearest,
    sky sentinel, no alpha in the colours) and at the demo's init camera (live variant, loader alpha 0xFF), plus smoothing."""
    for name, (rgb, hgt) in real_maps.items():
        mp = gpu_ctx.upload_map(rgb, hgt)                       # tools/png2data.py leaves alpha 0
        cam = fsb.Camera(512, 800, 78, 0, 100, 800, 1, SKY)     # tests/futspace.fut:129-136
        out = check(fsb, oracle, gpu_ctx, mp, rgb, hgt, cam, fsb.tests_variant_params(), 400, 800)
        assert len(np.unique(out)) >= 2, name                    # (on C7W/D7 this camera sits inside the terrain)
        mp.free()
        col = rgb | 0xFF000000                                   # c/freeimage_futspace.h:52
        mp = gpu_ctx.upload_map(col, hgt)
        cam = fsb.Camera(0.98, 0.6, 58, 2.2, 200, 800, 1.2, SKY)  # fut/interactive.fut:29-36
        cam.height = max(58.0, float(hgt[0, 0]))                # terrain_collision, :67-87
        for flags in (0, fsb.FLAG_SMOOTHING, fsb.FLAG_NO_TEXTURE):
            check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(flags=flags), 512, 640)
        mp.free()


def test_host_batch_double_buffering_and_launch_groups(fsb, oracle, gpu_ctx, fbm1024, monkeypatch):
    """fsb_render_batch splits a batch into 64 MB chunks copied out while the next chunk renders; the launch-group
    size (FSB_GROUP_POSES) splits a device batch.  Every frame must equal the single-frame render whatever the split."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    prm = fsb.default_params()
    h, w, n = 768, 1024, 50                      # 3 MiB frames: chunks of 21 poses -> three chunks, two staging buffers
    cams = camera_path(fsb, 1024, n, 600)
    batch = gpu_ctx.render_batch(cams, prm, mp, h, w)
    for i in (0, 20, 21, 22, 41, 42, 49):        # chunk boundaries
        assert np.array_equal(batch[i], gpu_ctx.render(cams[i], prm, mp, h, w)), i
    monkeypatch.setenv("FSB_GROUP_POSES", "7")   # read per call (fsb_api.c group_size)
    small = gpu_ctx.render_batch(cams[:20], prm, mp, 96, 128)
    monkeypatch.delenv("FSB_GROUP_POSES")
    ref = gpu_ctx.render_batch(cams[:20], prm, mp, 96, 128)
    assert np.array_equal(small, ref)
    mp.free()


def test_registered_host_frame_buffer(fsb, oracle, gpu_ctx, fbm1024):
    """fsb_host_register: the host's own frame buffer page-locked once (as a maintainer would do with lys's ctx->data),
    frames rendered into it repeatedly, then released."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    buf = np.zeros((300, 400), np.uint32)
    gpu_ctx.host_register(buf)
    for p in POSES[:3]:
        cam = fsb.Camera(*p, SKY)
        out = gpu_ctx.render(cam, fsb.default_params(), mp, 300, 400, out=buf)
        assert out is buf
        want = oracle.render(ocam(oracle, cam), oprm(oracle, fsb.default_params()), col, hgt & 0xFF, 300, 400)
        assert np.array_equal(buf, want)
    gpu_ctx.host_unregister(buf)
    with pytest.raises(fsb.FsbError):
        gpu_ctx.host_unregister(buf)           # not registered any more: reported, not fatal
    mp.free()


def test_one_pixel_wide_frame_with_horizon_below_the_frame(fsb, oracle, gpu_ctx, fbm1024):
    """Found by tools/soak_fuzz.py: w = 1 makes f32(w/2) = 0, inv_z is then 0 for every sample (row = horizon) except the
    z = 0 sample, whose inf * 0 = NaN converts to row 0 and fills the column; the occlusion bound must not skip it."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    for h, horizon, camh in ((27, 50.3, 453.0), (127, 158.8, 349.0), (64, 10.0, 300.0), (64, 64.0, 260.0)):
        for filt in (0, 1):
            for sentinel in (0, 1):
                cam = fsb.Camera(-174.69, 251.43, camh, 5.7278, horizon, 782.67, 1.113, SKY)
                check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(filter=filt, sentinel=sentinel), h, 1)
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(300.5, 200.25, 400, 1.0, 90.0, 500, 1.2, SKY),
          fsb.default_params(flags=fsb.FLAG_SMOOTHING), 40, 1)
    mp.free()


@pytest.mark.parametrize("slice_len,paint_seg", [(0, None), (1, None), (7, None), (32, None), (0, 0), (0, 1), (0, 3), (32, -1)],
                         ids=["slice0", "slice1", "slice7", "slice32", "paint_columns", "paint_seg1", "paint_seg3", "paint_default"])
def test_column_parallel_march_on_single_frames_and_colour_slices(fsb, oracle, gpu_ctx, fbm1024, monkeypatch, slice_len,
                                                                  paint_seg):
    """The column-parallel march is the batch path, followed either by the colour pass and expand as two launches
    (FSB_PAINT=0; FSB_COLOUR_SLICE cuts the colour pass into slices of the record lists) or by the paint kernel that does both
    (the default; FSB_PAINT_SEG = bands of 32 rows per warp, 0 = whole columns).  FSB_COLS_MIN_WARPS=0 forces the path for
    single frames.  Same frames as the oracle whatever the path: both filters, both sentinels, smoothing, ragged widths,
    ragged series in a batch, empty and one-sample series."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    monkeypatch.setenv("FSB_COLS_MIN_WARPS", "0")              # read per call (fsb_api.c)
    monkeypatch.setenv("FSB_COLOUR_SLICE", str(slice_len))
    if paint_seg is None:
        monkeypatch.setenv("FSB_PAINT", "0")
    elif paint_seg >= 0:
        monkeypatch.setenv("FSB_PAINT_SEG", str(paint_seg))
    n0 = gpu_ctx.launch_count
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*POSES[0], SKY), fsb.default_params(), 300, 417)
    assert gpu_ctx.launch_count - n0 == (4 if paint_seg is None else 3)   # set-up, march, colour + expand | paint
    for filt in (1, 0):
        for sentinel in (0, 1):
            prm = fsb.default_params(filter=filt, sentinel=sentinel)
            for p in POSES[:6]:
                check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 300, 417)
    prm = fsb.default_params(flags=fsb.FLAG_SMOOTHING)
    for p in POSES[:4]:
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 257, 95)
    prm = fsb.default_params()
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 40, 300, 1.2, SKY)
    for dist in (0.0004, 0.001, 0.6, 2.0, 30.0):            # n_z = 0, 1, 35, 63, 245
        cam.distance = dist
        check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, 64, 48)
    cams = camera_path(fsb, 1024, 5, 700)
    cams[3].distance = 90
    frames = gpu_ctx.render_batch(cams, prm, mp, 135, 240)
    for cam, got in zip(cams, frames):
        assert np.array_equal(got, oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt & 0xFF, 135, 240))
    # tests variant (z0 = 1, sky sentinel, nearest), a tall narrow frame, and the tallest frame the library takes
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(512, 800, 78, 0, 100, 800, 1, SKY), fsb.tests_variant_params(), 400, 33)
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(512.37, 512.73, 150, 2.2, 9000, 700, 1.2, SKY), fsb.default_params(), 32768, 40)
    mp.free()


@pytest.mark.parametrize("flags", [8, 8 | 16], ids=["default", "march_z"])
def test_smoothing_at_the_maximum_frame_height(fsb, oracle, gpu_ctx, fbm1024, monkeypatch, flags):
    """h = 32768 = 2^15: the neutral element (0, h, 0) of the smoothing scan does not fit the 15 row bits of a record;
    the first visible sample (k = 1 against the neutral k = 0) must still be blended exactly as the oracle does."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    for cols in ("0", None):
        if cols is not None:
            monkeypatch.setenv("FSB_COLS_MIN_WARPS", cols)
        for horizon in (9000, 31000):
            cam = fsb.Camera(512.37, 512.73, 150, 2.2, horizon, 700, 1.2, SKY)
            check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(flags=flags), 32768, 8)
    mp.free()


def test_batch_paths_agree(fsb, oracle, gpu_ctx, fbm1024, monkeypatch):
    """A batch large enough for the column-parallel march by default (80 poses x 60 groups of 32 columns): march + paint
    (the default), march + colour pass + expand (FSB_PAINT=0), the lanes-over-depth march (FSB_FLAG_MARCH_Z) and the oracle,
    frame by frame; launch counts tell the paths apart."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    cams = camera_path(fsb, 1024, 80, 900)
    for c in cams:
        c.horizon = 40
    h, w = 96, 1920
    p0 = gpu_ctx.launch_count
    c = gpu_ctx.render_batch(cams, fsb.default_params(), mp, h, w)
    monkeypatch.setenv("FSB_PAINT", "0")                       # read per call (fsb_api.c)
    n0 = gpu_ctx.launch_count
    a = gpu_ctx.render_batch(cams, fsb.default_params(), mp, h, w)
    n1 = gpu_ctx.launch_count
    b = gpu_ctx.render_batch(cams, fsb.default_params(flags=fsb.FLAG_MARCH_Z), mp, h, w)
    n2 = gpu_ctx.launch_count
    assert (n1 - n0) % 4 == 0 and (n2 - n1) % 3 == 0 and (n1 - n0) // 4 == (n2 - n1) // 3   # + the colour pass
    assert n0 - p0 == n2 - n1                                  # set-up, march, paint
    assert np.array_equal(a, b) and np.array_equal(a, c)
    for i in (0, 31, 32, 63, 79):
        want = oracle.render(ocam(oracle, cams[i]), oprm(oracle, fsb.default_params()), col, hgt & 0xFF, h, w)
        assert np.array_equal(a[i], want), i
    mp.free()


def test_paint_kernel_counters(fsb, oracle, gpu_ctx, fbm1024, monkeypatch):
    """The batch path's paint kernel reports its colour trips (fsb_context_paint_trips): every record is filtered in exactly
    one lane of one trip, so records <= 32 x trips, and the lanes are mostly busy (run-ahead through the ring)."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    monkeypatch.setenv("FSB_COLS_MIN_WARPS", "0")              # the host path renders in chunks of poses: keep them on the batch path
    cams = camera_path(fsb, 1024, 80, 900)
    for c in cams:
        c.horizon = 60
    gpu_ctx.set_profiling(True)
    try:
        gpu_ctx.get_counters()
        frames = gpu_ctx.render_batch(cams, fsb.default_params(), mp, 128, 1920)
        prof = gpu_ctx.get_profile()
        chunks, records = gpu_ctx.get_counters()
        trips = gpu_ctx.paint_trips
    finally:
        gpu_ctx.set_profiling(False)
    assert prof["colour"][0] < 0.25 * prof["expand"][0]        # no colour launch: the paint kernel sits in the expand slot
    assert records > 0 and trips > 0 and records <= 32 * trips
    assert records / (32.0 * trips) > 0.5
    want = oracle.render(ocam(oracle, cams[17]), oprm(oracle, fsb.default_params()), col, hgt & 0xFF, 128, 1920)
    assert np.array_equal(frames[17], want)
    mp.free()


def test_all_negative_terrain_integer_camera(fsb, oracle, gpu_ctx):
    """Bilinear sampling returns exactly 0 at integer coordinates (fact 9) -- above an all-negative unmasked terrain.
    With z0 = 0 the first sample sits at the camera: an integer camera position projects it to row 0 and blanks the
    column in the reference; the occlusion bound must not skip the chunk holding that sample."""
    rng = np.random.default_rng(9)
    q, r = 128, 256
    yy, xx = np.mgrid[0:q, 0:r]
    hgt = (-300 + 100 * np.sin(xx / 17.0) * np.cos(yy / 11.0)).astype(np.int32)     # all < -0.5
    col = rng.integers(0, 1 << 24, size=(q, r), dtype=np.uint64).astype(np.uint32) | 0xFF000000
    mp = gpu_ctx.upload_map(col, hgt, mask_heights=False)
    assert not mp.packed
    for cam_h in (-50.0, -250.0, 40.0):
        for x, y in ((100.0, 77.0), (100.5, 77.0), (100.25, 77.75)):
            for filt in (1, 0):
                cam = fsb.Camera(x, y, cam_h, 0.9, 60, 300, 1.2, SKY)
                a = check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, fsb.default_params(filter=filt), 120, 160, masked=False)
                b = gpu_ctx.render(cam, fsb.default_params(filter=filt, flags=fsb.FLAG_NO_CULL), mp, 120, 160)
                assert np.array_equal(a, b)
    mp.free()


def test_config5_16384_map_column_slabs_against_the_oracle(fsb, oracle, gpu_ctx):
    """BASELINE config 5 as stated: 7680x4320 frame, 16384^2 map, distance 4000.  Two of eight column slabs (device
    output at their offset in the frame) and one slab through fsb_render_columns (host output) against the oracle's
    render of the whole frame on the same map."""
    from futspace_b200.shard import column_bounds
    m, h, w = 16384, 4320, 7680
    col, hgt = fsb.terrain_fbm(m)
    mp = gpu_ctx.upload_map(col, hgt)
    cam = fsb.Camera(m / 2 + 0.37, m / 2 + 0.73, max(160.0, float(hgt[m // 2, m // 2]) + 20.0), 2.2, 0.3 * h, 4000, 1.2, SKY)
    prm = fsb.default_params()
    want = oracle.render(ocam(oracle, cam), oprm(oracle, prm), col, hgt, h, w)
    b = column_bounds(w, 8)
    dev = gpu_ctx.device_malloc(h * w * 4)
    for r in (0, 5):
        gpu_ctx.render_columns_device(cam, prm, mp, h, w, b[r], b[r + 1], dev + 4 * b[r], w)
    got = gpu_ctx.download(dev, (h, w))
    for r in (0, 5):
        assert np.array_equal(got[:, b[r]:b[r + 1]], want[:, b[r]:b[r + 1]]), r
    host = np.zeros((h, w), np.uint32)
    gpu_ctx.host_register(host)
    gpu_ctx.render_columns(cam, prm, mp, h, w, b[3], b[4], host.ctypes.data + 4 * b[3], w)
    gpu_ctx.host_unregister(host)
    assert np.array_equal(host[:, b[3]:b[4]], want[:, b[3]:b[4]])
    assert not host[:, :b[3]].any() and not host[:, b[4]:].any()      # nothing outside the slab was touched
    gpu_ctx.device_free(dev)
    mp.free()


@pytest.mark.parametrize("variant", ["one_warp_per_column", "four_warps_per_column", "split32", "split64",
                                     "four_warps_plain_expand", "one_warp_staged_expand_everywhere"])
def test_single_frame_march_variants(fsb, oracle, gpu_ctx, fbm1024, monkeypatch, variant):
    """Single frames and small batches on the texture path: four warps per column (fsb_march_frame.cu) below a size
    threshold, one warp per column (fsb_march_kernel) above it; FSB_FRAME_MAX_COLS moves the threshold; and behind
    FSB_SPLIT=1 the depth-parallel cluster march (fsb_march_split.cu; 32 or 64 depth segments per group of 32 columns,
    FSB_SPLIT_WARPS); the expand behind them with and without the band's records staged in shared memory
    (FSB_EXPAND_STAGE).  All against the oracle: filters, sentinels, smoothing, full evaluation, ragged and degenerate
    sizes, short and long series (the longest falls back from the split march), batches."""
    col, hgt = fbm1024
    mp = gpu_ctx.upload_map(col, hgt)
    if variant.startswith("split"):
        monkeypatch.setenv("FSB_SPLIT", "1")
        monkeypatch.setenv("FSB_SPLIT_WARPS", variant[5:])
        n0 = gpu_ctx.launch_count
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*POSES[0], SKY), fsb.default_params(), 300, 417)
        assert gpu_ctx.launch_count - n0 == 3      # march, colour, expand: no set-up launch
    else:
        monkeypatch.setenv("FSB_FRAME_MAX_COLS", "0" if variant.startswith("one_warp") else "100000000")
        # the expand of single frames stages a band's records in shared memory (fsb_expand4s_kernel); batches walk the lists
        if variant.endswith("plain_expand"):
            monkeypatch.setenv("FSB_EXPAND_STAGE", "0")
        elif variant.endswith("everywhere"):
            monkeypatch.setenv("FSB_EXPAND_STAGE", "1")
    for filt in (1, 0):
        for sentinel in (0, 1):
            for flags in (0, fsb.FLAG_NO_CULL):
                prm = fsb.default_params(filter=filt, sentinel=sentinel, flags=flags)
                for p in POSES[:6]:
                    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), prm, 300, 417)
    for p in POSES[:4]:
        check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(*p, SKY), fsb.default_params(flags=fsb.FLAG_SMOOTHING), 257, 95)
    prm = fsb.default_params()
    cam = fsb.Camera(512.37, 512.73, 180, 2.2, 40, 300, 1.2, SKY)
    for h, w in ((1, 1), (7, 5), (33, 9), (257, 8)):
        check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, h, w)
    for dist in (0.0004, 0.001, 0.6, 2.0, 5.0, 30.0, 2000.0, 7000.0, 60000.0):   # n_z = 0, 1, 35, 63, 100, 245, 2000, 3742, 10954
        cam.distance = dist
        check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, 64, 48)
    for cam_h in (400, 700, 5000):                                  # camera above the terrain: the skipped prefix
        cam = fsb.Camera(512.37, 512.73, cam_h, 2.2, 20, 900, 1.2, SKY)
        check(fsb, oracle, gpu_ctx, mp, col, hgt, cam, prm, 200, 333)
    for n, dist in ((3, 700), (7, 1500)):                           # small batches, ragged series
        cams = camera_path(fsb, 1024, n, dist)
        cams[n // 2].distance = 90
        frames = gpu_ctx.render_batch(cams, prm, mp, 135, 240)
        for c, got in zip(cams, frames):
            assert np.array_equal(got, oracle.render(ocam(oracle, c), oprm(oracle, prm), col, hgt & 0xFF, 135, 240))
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(512.37, 512.73, 150, 2.2, 9000, 700, 1.2, SKY), fsb.default_params(), 32768, 40)
    check(fsb, oracle, gpu_ctx, mp, col, hgt, fsb.Camera(512, 800, 78, 0, 100, 800, 1, SKY), fsb.tests_variant_params(), 400, 800)
    mp.free()


def test_local_occlusion_bound_changes_no_pixel(fsb, oracle, gpu_ctx, monkeypatch):
    """The column-parallel march skips a chunk of 32 depth steps x 32 columns when the pyramid of local height maxima
    (built at map upload for power-of-two packed maps) proves that no sample of it can pass `occlude`.  Frames must be
    those of the oracle with and without it (FSB_LOCAL_CULL=0), on rugged and flat terrain, both filters, both
    sentinels, map sizes down to one pyramid level, cameras above / below / inside the terrain and far outside the map;
    the counter of evaluated chunks must drop."""
    monkeypatch.setenv("FSB_COLS_MIN_WARPS", "0")          # the column-parallel march also for single frames
    rng = np.random.default_rng(11)
    col, hgt = fsb.terrain_fbm(1024)
    maps = [(col, hgt)]
    yy, xx = np.mgrid[0:64, 0:256]
    maps.append((rng.integers(0, 1 << 24, size=(64, 256), dtype=np.uint64).astype(np.uint32) | 0xFF000000,
                 (128 + 120 * np.sin(xx / 9.0) * np.cos(yy / 5.0)).astype(np.int32)))      # non-square, steep
    maps.append((rng.integers(0, 1 << 24, size=(2, 4), dtype=np.uint64).astype(np.uint32) | 0xFF000000,
                 np.array([[0, 255, 3, 9], [200, 1, 77, 30]], np.int32)))                   # one pyramid level
    spikes = np.zeros((512, 512), np.int32)
    spikes[rng.integers(0, 512, 300), rng.integers(0, 512, 300)] = 255                        # isolated one-texel spikes
    maps.append((rng.integers(0, 1 << 24, size=(512, 512), dtype=np.uint64).astype(np.uint32) | 0xFF000000, spikes))
    gpu_ctx.set_profiling(True)
    saved = []
    for color, height in maps:
        mp = gpu_ctx.upload_map(color, height)
        q, r = height.shape
        for filt in (1, 0):
            for sentinel in (0, 1):
                prm = fsb.default_params(filter=filt, sentinel=sentinel)
                for (x, y, ch, ang, hor, dist, fov) in POSES[:6] + [(r / 2 + 0.3, q / 2 + 0.6, 20, 5.1, 150, 1500, 1.2),
                                                                    (r / 3 + 0.3, q / 3 + 0.6, 300, 0.4, 60, 2500, 0.7)]:
                    cam = fsb.Camera(x, y, ch, ang, hor, dist, fov, SKY)
                    gpu_ctx.get_counters()
                    a = check(fsb, oracle, gpu_ctx, mp, color, height, cam, prm, 200, 333)
                    with_cull = gpu_ctx.get_counters()[0]
                    monkeypatch.setenv("FSB_LOCAL_CULL", "0")
                    b = gpu_ctx.render(cam, prm, mp, 200, 333)
                    monkeypatch.delenv("FSB_LOCAL_CULL")
                    without = gpu_ctx.get_counters()[0]
                    assert np.array_equal(a, b)
                    assert with_cull <= without
                    saved.append((with_cull, without))
        mp.free()
    gpu_ctx.set_profiling(False)
    assert sum(a for a, _ in saved) < 0.9 * sum(b for _, b in saved)
