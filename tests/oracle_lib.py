"""ctypes binding for the parity oracle (oracle/libfs_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(ORACLE_DIR, "libfs_oracle.so")

F_NEAREST, F_BILINEAR = 0, 1
S_ZERO, S_SKY = 0, 1
F2I_SATURATE, F2I_X86, F2I_MODERN = 0, 1, 2


class Camera(ctypes.Structure):
    """fso_camera == fut/voxel_renderer.fut:5-12"""
    _fields_ = [(n, ctypes.c_float) for n in "x y height angle horizon distance fov".split()] + [
        ("sky_color", ctypes.c_uint32)]


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in "z0 delta invz_param1 invz_param2".split()] + [
        (n, ctypes.c_int32) for n in "filter sentinel f2i_mode smoothing".split()]


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("fs_oracle.c", "fs_oracle.h", "Makefile")]
    src.append(os.path.join(os.path.dirname(ORACLE_DIR), "futspace_b200", "csrc", "fsb_terrain.c"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, ci, cf, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint32
        P = ctypes.POINTER
        L.fso_params_default.argtypes = [P(Params)]
        L.fso_params_tests_variant.argtypes = [P(Params)]
        L.fso_get_zs.argtypes = [cf, cf, cf, vp, ci]
        L.fso_get_zs.restype = ci
        L.fso_mix.argtypes = [cf, u32, cf, u32]
        L.fso_mix.restype = u32
        L.fso_scale.argtypes = [u32, cf]
        L.fso_scale.restype = u32
        for name, rt in (("fso_height_nearest", cf), ("fso_height_bilinear", cf),
                         ("fso_color_nearest", u32), ("fso_color_bilinear", u32)):
            f = getattr(L, name)
            f.argtypes = [vp, ci, ci, cf, cf, ci]
            f.restype = rt
        L.fso_render.argtypes = [P(Camera), P(Params), vp, vp, ci, ci, ci, ci, vp, ci, ci]
        L.fso_render.restype = ci
        L.fso_render_split.argtypes = [P(Camera), P(Params), vp, ci, ci, vp, ci, ci, ci, ci, vp, ci, ci]
        L.fso_render_split.restype = ci
        L.fso_render_literal.argtypes = [P(Camera), P(Params), vp, vp, ci, ci, ci, ci, vp]
        L.fso_render_literal.restype = ci
        L.fso_mask_heights.argtypes = [vp, ctypes.c_long]
        L.fso_bake_shadows.argtypes = [vp, vp, ci, ci, P(cf), ci, ci, vp, ci]
        L.fso_sun_vector.argtypes = [cf, cf, P(cf)]
        L.fso_interpolate.argtypes = [ci, vp, ci, ci, vp]
        L.fso_interpolate.restype = ci
        L.fso_interpolate2.argtypes = [vp, ci, ci, vp]
        L.fsb_terrain_fbm.argtypes = [ci, ctypes.c_uint64, vp, vp]
        L.fsb_terrain_fbm.restype = ci
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().fso_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def tests_variant_params(**kw):
    p = Params()
    lib().fso_params_tests_variant(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def get_zs(delta, dist, z0, cap=1 << 20):
    buf = np.zeros(cap, np.float32)
    n = lib().fso_get_zs(delta, dist, z0, buf.ctypes.data, cap)
    if n < 0:
        raise ValueError("invalid z-series arguments")
    return buf[:n].copy()


def terrain_fbm(m, seed=0x5EED5EED):
    """The benchmark terrain, from the generator source compiled into the oracle library: the CPU arm
    of bench.py needs no product library."""
    color = np.empty((m, m), np.uint32)
    height = np.empty((m, m), np.int32)
    if lib().fsb_terrain_fbm(m, seed, color.ctypes.data, height.ctypes.data):
        raise ValueError("terrain_fbm(m=%d)" % m)
    return color, height


def _check_maps(color, height):
    color = np.ascontiguousarray(color, dtype=np.uint32)
    height = np.ascontiguousarray(height, dtype=np.int32)
    assert color.ndim == 2 and color.shape == height.shape
    return color, height


def render(cam, prm, color, height, h, w, eval_all_colors=False, nthreads=0):
    color, height = _check_maps(color, height)
    out = np.empty((h, w), np.uint32)
    rc = lib().fso_render(ctypes.byref(cam), ctypes.byref(prm), color.ctypes.data, height.ctypes.data,
                          color.shape[0], color.shape[1], h, w, out.ctypes.data,
                          1 if eval_all_colors else 0, nthreads)
    if rc:
        raise RuntimeError("fso_render failed rc=%d" % rc)
    return out


def render_split(cam, prm, color, height, h, w, nthreads=0):
    """Colour and height maps of different sizes (the state update_map leaves for maps that are not 1024 x 1024)."""
    color = np.ascontiguousarray(color, dtype=np.uint32)
    height = np.ascontiguousarray(height, dtype=np.int32)
    out = np.empty((h, w), np.uint32)
    rc = lib().fso_render_split(ctypes.byref(cam), ctypes.byref(prm), color.ctypes.data, color.shape[0], color.shape[1],
                                height.ctypes.data, height.shape[0], height.shape[1], h, w, out.ctypes.data, 0, nthreads)
    if rc:
        raise RuntimeError("fso_render_split failed rc=%d" % rc)
    return out


def render_literal(cam, prm, color, height, h, w):
    color, height = _check_maps(color, height)
    out = np.empty((h, w), np.uint32)
    rc = lib().fso_render_literal(ctypes.byref(cam), ctypes.byref(prm), color.ctypes.data,
                                  height.ctypes.data, color.shape[0], color.shape[1], h, w, out.ctypes.data)
    if rc:
        raise RuntimeError("fso_render_literal failed rc=%d" % rc)
    return out


def sun_vector(sun_height, sun_ang):
    v = (ctypes.c_float * 3)()
    lib().fso_sun_vector(sun_height, sun_ang, v)
    return [v[0], v[1], v[2]]


def bake_shadows(color, height, sun, out_q=None, out_r=None, nthreads=0):
    color, height = _check_maps(color, height)
    q, r = color.shape
    out_q, out_r = out_q or q, out_r or r
    out = np.empty((out_q, out_r), np.uint32)
    s = (ctypes.c_float * 3)(*sun)
    lib().fso_bake_shadows(color.ctypes.data, height.ctypes.data, q, r, s, out_q, out_r, out.ctypes.data, nthreads)
    return out


def interpolate(pd, img):
    img = np.ascontiguousarray(img, np.uint32)
    out = np.empty_like(img)
    if lib().fso_interpolate(pd, img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data):
        raise ValueError("interpolate needs h <= w (fut/effects.fut:36-41 wraps x with % h)")
    return out


def interpolate2(img):
    img = np.ascontiguousarray(img, np.uint32)
    out = np.empty_like(img)
    lib().fso_interpolate2(img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data)
    return out
