"""CPU tests of the oracle itself: the reference's only known-answer relation (z-series), the
documented properties of SURVEY.md facts 8/9, two independent formulations against each other,
and the committed golden frames."""
import ctypes
import math

import numpy as np
import pytest

from conftest import SKY


def test_zs_known_answer_tests_constants(oracle):
    # tests/solve_arithm.fut:1-8 vs tests/javascript_zs.py:2-8 : get_zs(0.005, 800, 1)
    zs = oracle.get_zs(0.005, 800.0, 1.0)
    assert len(zs) == 400
    js, z, dz = [], 1.0, 1.0
    while z <= 800:
        js.append(z)
        dz += 0.005
        z += dz
    assert len(js) == 400
    np.testing.assert_allclose(zs, np.array(js), rtol=2e-6)
    assert zs[0] == np.float32(1.0) and zs[1] == np.float32(2.005) and zs[3] == np.float32(4.03)
    assert zs[2] == np.float32(3.0149999)


@pytest.mark.parametrize("dist,n", [(800, 1265), (1000, 1414), (2000, 2000), (4000, 2828)])
def test_zs_live_constants(oracle, dist, n):
    # fut/voxel_renderer.fut:103-108 : z0 = 0, d = 0.001 ; SURVEY.md 8 sizes
    zs = oracle.get_zs(0.001, float(dist), 0.0)
    assert len(zs) == n
    assert zs[0] == 0.0  # the z = 0 trap (fact 8)
    assert np.all(np.diff(zs) > 0) and zs[-1] <= dist


@pytest.mark.parametrize("dist,n", [(1000, 463), (2000, 716), (4000, 1081)])
def test_zs_tests_constants(oracle, dist, n):
    assert len(oracle.get_zs(0.005, float(dist), 1.0)) == n


def test_zs_invalid(oracle):
    with pytest.raises(ValueError):
        oracle.get_zs(0.001, -5000.0, 0.0)


def test_mix_basics(oracle):
    L = oracle.lib()
    c1, c2 = 0xFF102030, 0xFF908070
    assert L.fso_mix(1.0, c1, 0.0, c2) == c1          # all weight on one side is the identity on 8-bit colours
    assert L.fso_mix(0.0, c1, 1.0, c2) == c2
    assert L.fso_mix(0.0, c1, 0.0, c2) == 0            # 0/0 -> NaN -> 0
    assert L.fso_mix(0.5, c1, 0.5, c1) == c1
    m = L.fso_mix(0.25, 0xFF000000, 0.75, 0xFFFFFFFF)  # sqrt(0.75) * 255 = 220.8 -> 220
    assert m == 0xFFDCDCDC
    assert L.fso_scale(0xFF9090E0, 0.1) == 0x190E0E16  # fut/interactive.fut:163 at sun_height 0.1


def test_bilinear_integer_coordinate_degenerates(oracle, c1w_d1):
    rgb, hgt = c1w_d1
    col = np.ascontiguousarray(rgb | 0xFF000000, np.uint32)
    hgt = np.ascontiguousarray(hgt)
    L = oracle.lib()
    # floor == ceil -> both weights 0 -> height 0 and colour 0 (fact 9)
    assert L.fso_height_bilinear(hgt.ctypes.data, 1024, 1024, 10.0, 20.5, 0) == 0.0
    assert L.fso_color_bilinear(col.ctypes.data, 1024, 1024, 10.0, 20.5, 0) == 0
    assert L.fso_color_bilinear(col.ctypes.data, 1024, 1024, 10.5, 20.0, 0) == 0
    # non-integer: alpha stays 0xFF, wraps with floored modulo
    c = L.fso_color_bilinear(col.ctypes.data, 1024, 1024, -0.5, 1023.5, 0)
    assert c >> 24 == 0xFF
    h = L.fso_height_bilinear(hgt.ctypes.data, 1024, 1024, -0.5, 1023.5, 0)
    expect = 0.25 * (hgt[1023, 1023] + hgt[1023, 0] + hgt[0, 1023] + hgt[0, 0])
    assert h == pytest.approx(expect)


def test_nearest_truncates_toward_zero(oracle, c1w_d1):
    rgb, hgt = c1w_d1
    hgt = np.ascontiguousarray(hgt)
    L = oracle.lib()
    # i32.f32 truncates: -0.5 -> 0 (not -1)   fut/render_functions.fut:63-64
    assert L.fso_height_nearest(hgt.ctypes.data, 1024, 1024, -0.5, -0.5, 0) == float(hgt[0, 0])
    assert L.fso_height_nearest(hgt.ctypes.data, 1024, 1024, -1.5, 2.5, 0) == float(hgt[2, 1023])


def test_golden_tests_variant(oracle, c1w_d1, golden_frames):
    rgb, hgt = c1w_d1
    cam = oracle.Camera(512, 800, 78, 0, 100, 800, 1, SKY)   # tests/futspace.fut:129-136
    out = oracle.render(cam, oracle.tests_variant_params(), rgb, hgt, 400, 800)
    assert np.array_equal(out, golden_frames["tests_variant_400x800"])
    assert abs((out == SKY).mean() - 0.2258) < 1e-4 and len(np.unique(out)) == 102   # SURVEY.md 8c smoke values
    lit = oracle.render_literal(cam, oracle.tests_variant_params(), rgb, hgt, 400, 800)
    assert np.array_equal(out, lit)


def test_golden_live_variant(oracle, c1w_d1, golden_frames):
    rgb, hgt = c1w_d1
    cam = oracle.Camera(0.98, 0.6, 58, 2.2, 200, 1000, 1.2, SKY)  # fut/interactive.fut:29-36, distance of config 1
    out = oracle.render(cam, oracle.default_params(), rgb | 0xFF000000, hgt, 768, 1024)
    assert np.array_equal(out, golden_frames["live_init_768x1024_d1000"])


@pytest.mark.parametrize("filt", [0, 1])
@pytest.mark.parametrize("sentinel", [0, 1])
def test_sequential_equals_literal_pipeline(oracle, fbm1024, filt, sentinel):
    # the march restatement vs the reference's scan/scatter/scan taken literally (voxel_renderer.fut:229-250)
    col, hgt = fbm1024
    prm = oracle.default_params(filter=filt, sentinel=sentinel)
    for cam in (oracle.Camera(512.37, 512.73, 180, 2.2, 60, 300, 1.2, SKY),
                oracle.Camera(100.0, 7.0, 30, -0.7, 150, 200, 0.8, SKY),      # integer coords, below terrain
                oracle.Camera(3.25, 900.5, 260, 4.0, -20, 250, 1.5, SKY)):    # horizon < 0
        a = oracle.render(cam, prm, col, hgt, 200, 320)
        b = oracle.render_literal(cam, prm, col, hgt, 200, 320)
        c = oracle.render(cam, prm, col, hgt, 200, 320, eval_all_colors=True, nthreads=2)
        assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("mode", [1, 2])
def test_flat_frame_under_cpu_float_to_int(oracle, c1w_d1, mode):
    # fact 8: with z0 = 0 the CPU conversions make every column one flat colour
    rgb, hgt = c1w_d1
    cam = oracle.Camera(0.98, 0.6, 58, 2.2, 40, 300, 1.2, SKY)
    out = oracle.render(cam, oracle.default_params(f2i_mode=mode), rgb | 0xFF000000, hgt, 128, 160)
    assert all(len(np.unique(out[:, j])) == 1 for j in range(out.shape[1]))
    sat = oracle.render(cam, oracle.default_params(), rgb | 0xFF000000, hgt, 128, 160)
    assert len(np.unique(sat)) > 100


def test_f2i_modes_agree_when_z0_positive(oracle, c1w_d1):
    rgb, hgt = c1w_d1
    cam = oracle.Camera(512, 800, 78, 0, 100, 400, 1, SKY)
    frames = [oracle.render(cam, oracle.tests_variant_params(f2i_mode=m), rgb, hgt, 100, 200) for m in (0, 1, 2)]
    assert np.array_equal(frames[0], frames[1]) and np.array_equal(frames[0], frames[2])


def test_degenerate_sizes(oracle, fbm1024):
    col, hgt = fbm1024
    cam = oracle.Camera(10.3, 20.7, 300, 1.0, 5, 0.0004, 1.2, SKY)      # n_z = 0 -> all sky
    out = oracle.render(cam, oracle.default_params(), col, hgt, 7, 5)
    assert (out == SKY).all()
    cam.distance = 0.002                                                  # n_z = 2 (z = 0, 0.001)
    assert len(oracle.get_zs(0.001, 0.002, 0.0)) == 2
    out = oracle.render(cam, oracle.default_params(), col, hgt, 1, 1)
    assert out.shape == (1, 1)


def test_effects_post_passes_oracle(oracle):
    img = np.full((6, 8), 0xFF808080, np.uint32)
    assert np.array_equal(oracle.interpolate2(img), img)          # mixing equal colours is the identity
    assert np.array_equal(oracle.interpolate(2, img), img)
    img[2, 3] = 0xFFFFFFFF
    out = oracle.interpolate2(img)
    assert out[2, 3] != img[2, 3] and out[2, 2] != img[2, 2] and out[0, 0] == img[0, 0]
    with pytest.raises(ValueError):
        oracle.interpolate(1, np.zeros((8, 6), np.uint32))


def test_smoothing_on_two_formulations_agree(oracle, c1w_d1):
    """Smoothing #on (fut/voxel_renderer.fut:175-213): the sequential statement (list of lowering samples, spans
    blended when three consecutive samples lower the y-buffer) against the reference pipeline taken literally
    (scan occlude2 / rotate / scatter in index order / scan fill_vline3 / map), and against #off."""
    rgb, hgt = c1w_d1
    col = rgb | np.uint32(0xFF000000)
    changed = 0
    for cam in (oracle.Camera(512.3, 800.7, 78, 0.3, 100, 300, 1.0, SKY), oracle.Camera(300.5, 200.25, 120, 2.2, 60, 500, 1.2, SKY),
                oracle.Camera(100.0, 7.0, 30, -0.7, 150, 400, 0.8, SKY)):
        for filt in (0, 1):
            on = oracle.default_params(smoothing=1, filter=filt)
            a = oracle.render(cam, on, col, hgt, 160, 200)
            b = oracle.render_literal(cam, on, col, hgt, 160, 200)
            assert np.array_equal(a, b)
            off = oracle.render(cam, oracle.default_params(filter=filt), col, hgt, 160, 200)
            assert (a != off).mean() < 0.5
            changed += int((a != off).any())
            if filt == 0:   # nearest C1W colours are never 0: same silhouette
                assert np.array_equal(a == SKY, off == SKY)
    assert changed >= 4
    with pytest.raises(RuntimeError):   # no smoothing in the sky-sentinel renderer (fut/voxel_renderer_new.fut)
        oracle.render(cam, oracle.default_params(smoothing=1, sentinel=1), col, hgt, 16, 16)


def test_real_map_pairs_two_formulations(oracle, real_maps):
    """The sequential march and the literal scan / scatter / scan pipeline agree on every committed real map pair."""
    for name, (rgb, hgt) in real_maps.items():
        cam = oracle.Camera(512, 800, 78, 0, 100, 300, 1, SKY)
        for prm in (oracle.tests_variant_params(), oracle.default_params(), oracle.default_params(smoothing=1)):
            col = rgb if prm.sentinel else rgb | np.uint32(0xFF000000)
            a = oracle.render(cam, prm, col, hgt, 100, 160)
            b = oracle.render_literal(cam, prm, col, hgt, 100, 160)
            assert np.array_equal(a, b), name
