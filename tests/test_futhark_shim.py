"""The generated-API names (include/libfutspace.h): host-side state machine of fut/interactive.fut on CPU,
and the full c/interactive.c call sequence on the GPU against the oracle."""
import math

import ctypes

import numpy as np
import pytest

from conftest import SKY
import futhark_shim as FS

f32 = np.float32


@pytest.fixture()
def session(fsb):
    s = FS.Session()
    s.error()          # on a box without a GPU the context carries an error; the state machine still works
    s.init()
    yield s
    s.close()


def test_init_state(session):
    # fut/interactive.fut:29-36,54-55 via text_content (:185-186)
    x, y, angle, height, horizon, distance, sun_h, sun_a, fov = session.text_content()
    assert (f32(x), f32(y), height, f32(angle), horizon, distance, f32(fov)) == (f32(0.98), f32(0.6), 58.0, f32(2.2), 200.0, 800.0, f32(1.2))
    assert (f32(sun_h), f32(sun_a)) == (f32(0.1), f32(0.1))


def test_keys_and_step(session):
    # process_inputs, fut/interactive.fut:89-125: every update reads the state before the step
    session.key(True, ord("w"))
    session.key(True, ord("a"))
    session.key(True, ord("q"))
    session.key(True, FS.SDLK_UP)
    session.key(True, ord("o"))
    session.step()
    x, y, angle, height, horizon, distance, sun_h, sun_a, fov = session.text_content()
    a0 = f32(2.2)
    assert f32(x) == f32(f32(0.98) - f32(3) * f32(math.sin(a0))) or abs(x - (0.98 - 3 * math.sin(2.2))) < 1e-6
    assert abs(y - (0.6 - 3 * math.cos(2.2))) < 1e-6
    assert f32(angle) == f32(a0 + f32(0.10))
    assert horizon == 220.0 and distance == 830.0
    assert f32(fov) == f32(f32(1.2) + f32(0.1))
    # keyup stops the motion; opposite keys
    for k in (ord("w"), ord("a"), ord("q"), FS.SDLK_UP, ord("o")):
        session.key(False, k)
    before = session.text_content()
    session.step()
    assert session.text_content() == before
    session.key(True, ord("s")); session.key(True, ord("d")); session.key(True, ord("e")); session.key(True, FS.SDLK_DOWN)
    session.key(True, ord("l")); session.key(True, ord("r"))
    session.step()
    after = session.text_content()
    assert f32(after[2]) == f32(f32(before[2]) - f32(0.10)) and after[4] == before[4] - 20 and after[5] == before[5] - 30
    assert after[3] == before[3] + 10


def test_terrain_collision_and_misc_events(session):
    # no map yet: altitude is the dummy [[0],[0]] (fut/interactive.fut:38-43): height is clamped at 0
    session.key(True, ord("f"))
    for _ in range(8):
        session.step()
    assert session.text_content()[3] == 0.0      # 58 - 6*10 would be negative
    session.mouse()
    session.resize(480, 640)
    assert session.text_content()[3] == 0.0


def test_unknown_key_is_ignored(session):
    before = session.text_content()
    session.key(True, 0x7F)
    session.step()
    assert session.text_content() == before


@pytest.mark.gpu
def test_interactive_c_call_sequence_matches_oracle(fsb, oracle, c1w_d1):
    rgb, hgt = c1w_d1
    col = rgb | 0xFF000000                         # c/freeimage_futspace.h:52
    raw_h = hgt | 0x00ABCD00                       # junk above bit 7: update_map masks with & 0xFF (fut/interactive.fut:189)
    s = FS.Session()
    assert s.error() is None
    s.init()
    s.update_map(col, raw_h)                       # load_map, c/interactive.c:25-57
    s.resize(384, 512)
    s.step()                                       # sets cam.sky_color = argb.scale 0xFF9090e0 sun_height (:163)
    frame = s.render()
    sun = oracle.sun_vector(0.1, 0.1)
    shadowed = oracle.bake_shadows(col, hgt, sun, 1024, 1024)
    sky = oracle.lib().fso_scale(0xFF9090E0, 0.1)
    assert sky == 0x190E0E16
    cam = oracle.Camera(0.98, 0.6, 58, 2.2, 200, 800, 1.2, sky)   # terrain under (0,0)... stays 58 unless higher
    ground = float(hgt[0, 0])
    cam.height = max(58.0, ground)
    want = oracle.render(cam, oracle.default_params(), shadowed, hgt, 384, 512)
    assert np.array_equal(frame, want)
    # walk and turn for a few frames, raise the sun (re-bakes the shadows), compare again
    s.key(True, ord("w")); s.key(True, ord("a")); s.key(True, ord("j"))
    for _ in range(3):
        s.step()
    x, y, angle, height, horizon, distance, sun_h, sun_a, fov = s.text_content()
    frame = s.render()
    # the shadow map and sky in use come from the sun angles BEFORE the last step (fut/interactive.fut:126-136,163)
    v2 = f32(f32(0.1) - f32(0.005)) - f32(0.005)
    assert f32(sun_h) == f32(v2 - f32(0.005))
    sun = oracle.sun_vector(float(v2), sun_a)
    shadowed = oracle.bake_shadows(col, hgt, sun, 1024, 1024)
    sky = oracle.lib().fso_scale(0xFF9090E0, float(v2))
    cam = oracle.Camera(x, y, height, angle, horizon, distance, fov, sky)
    want = oracle.render(cam, oracle.default_params(), shadowed, hgt, 384, 512)
    assert np.array_equal(frame, want)
    # key `2` toggles smoothing (fut/interactive.fut:153-159) on the step after the key-down event
    s.key(False, ord("w")); s.key(False, ord("a")); s.key(False, ord("j"))
    s.key(True, ord("2"))
    s.step()
    s.key(False, ord("2"))
    s.step()
    x, y, angle, height, horizon, distance, sun_h2, sun_a2, fov = s.text_content()
    assert (sun_h2, sun_a2) == (sun_h, sun_a)      # no sun key held: the shadow map baked above stays in use
    cam = oracle.Camera(x, y, height, angle, horizon, distance, fov, oracle.lib().fso_scale(0xFF9090E0, sun_h))
    frame = s.render()
    want = oracle.render(cam, oracle.default_params(smoothing=1), shadowed, hgt, 384, 512)
    assert np.array_equal(frame, want)
    assert not np.array_equal(frame, oracle.render(cam, oracle.default_params(), shadowed, hgt, 384, 512))
    s.close()


@pytest.mark.gpu
def test_render_before_update_map_draws_the_dummy_landscape(fsb, oracle):
    """init's placeholder landscape is altitude = color = shadowed_color = [[0],[0]] (fut/interactive.fut:38-43): every
    sample has height 0 and the empty colour, so `render` returns the sky colour everywhere -- 0 before the first step,
    argb.scale 0xFF9090e0 sun_height after it (:163)."""
    s = FS.Session()
    s.init()
    s.resize(96, 128)
    frame = s.render()
    assert frame.shape == (96, 128) and (frame == 0).all()
    s.step()
    frame = s.render()
    sky = oracle.lib().fso_scale(0xFF9090E0, 0.1)
    assert (frame == sky).all()
    x, y, angle, height, horizon, distance, _, _, fov = s.text_content()
    want = oracle.render(oracle.Camera(x, y, height, angle, horizon, distance, fov, sky), oracle.default_params(),
                         np.zeros((2, 1), np.uint32), np.zeros((2, 1), np.int32), 96, 128)
    assert np.array_equal(frame, want)
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(300, 517), (2048, 1024)])
def test_update_map_of_a_map_that_is_not_1024_square(fsb, oracle, fbm1024, shape):
    """update_map bakes lsc.shadowed_color at 1024 x 1024 whatever the map size (fut/effects.fut:124-125) and keeps
    lsc.altitude at the map's size (fut/interactive.fut:188-198); `render` then wraps each by its own size."""
    col, hgt = fbm1024
    q, r = shape
    cmap = np.ascontiguousarray(np.tile(col, (2, 1))[:q, :r])
    hmap = np.ascontiguousarray(np.tile(hgt, (2, 1))[:q, :r])
    s = FS.Session()
    s.init()
    s.update_map(cmap, hmap)
    s.resize(240, 320)
    s.step()
    frame = s.render()
    sun = oracle.sun_vector(0.1, 0.1)
    shadowed = oracle.bake_shadows(cmap, hmap, sun, 1024, 1024)
    sky = oracle.lib().fso_scale(0xFF9090E0, 0.1)
    x, y, angle, height, horizon, distance, _, _, fov = s.text_content()
    cam = oracle.Camera(x, y, height, angle, horizon, distance, fov, sky)
    want = oracle.render_split(cam, oracle.default_params(), shadowed, hmap, 240, 320)
    assert np.array_equal(frame, want)
    assert len(np.unique(frame)) > 20
    s.close()


@pytest.mark.gpu
def test_values_into_a_registered_buffer_needs_no_staging_copy(fsb, oracle, c1w_d1):
    """futhark_values_u32_2d into a page-locked destination (fsb_host_register on the host's frame buffer) is a
    single DMA; into pageable memory it goes through the shim's pinned staging buffer.  Same pixels either way."""
    rgb, hgt = c1w_d1
    s = FS.Session()
    s.init()
    s.update_map(rgb | 0xFF000000, hgt)
    s.resize(200, 256)
    s.step()
    plain = s.render()
    L = s.L
    out = FS.vp()
    assert L.futhark_entry_render(s.ctx, ctypes.byref(out), s.state) == 0
    buf = np.zeros((200, 256), np.uint32)
    raw = ctypes.c_void_p.from_address(s.ctx)   # struct futhark_context { fsb_context *fsb; ... }
    F = fsb.lib()
    assert F.fsb_host_register(raw, buf.ctypes.data, buf.nbytes) == 0
    assert F.fsb_host_is_registered(raw, buf.ctypes.data) == 1 and F.fsb_host_is_registered(raw, plain.ctypes.data) == 0
    assert L.futhark_values_u32_2d(s.ctx, out, buf.ctypes.data) == 0
    assert L.futhark_context_sync(s.ctx) == 0
    assert F.fsb_host_unregister(raw, buf.ctypes.data) == 0
    L.futhark_free_u32_2d(s.ctx, out)
    assert np.array_equal(buf, plain)
    s.close()
