"""futspace_b200 -- ctypes binding of the C-ABI in include/futspace_b200.h.

The product is the shared library futspace_b200/libfutspace_b200.so (hand-written sm_100a CUDA
kernels + a C host layer).  This module is only the thin Python view of that C-ABI which the
tests and bench.py use; it contains no rendering logic and no CPU fallback: if the library is
missing, or no sm_100 GPU is present, constructing a Context raises.

Reference interface mirrored (see include/futspace_b200.h for the per-function mapping):
`futhark_context_new`, `futhark_new_u32_2d`/`futhark_new_i32_2d` + `futhark_entry_update_map`,
`futhark_entry_render`, `futhark_values_u32_2d`, `futhark_context_sync`
(c/interactive.c:50-56,99,144-148; fut/interactive_entrypoints.fut:19-32).
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfutspace_b200.so")
CSRC = os.path.join(_PKG, "csrc")

FILTER_NEAREST, FILTER_BILINEAR = 0, 1
SENTINEL_ZERO, SENTINEL_SKY = 0, 1
F2I_SATURATE, F2I_X86, F2I_MODERN = 0, 1, 2
FLAG_FORCE_GENERIC = 1
FLAG_NO_TEXTURE = 2
FLAG_NO_CULL = 4
FLAG_SMOOTHING = 8
FLAG_MARCH_Z = 16

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_RANGE, ERR_NO_DEVICE = range(6)


class Camera(ctypes.Structure):
    """fsb_camera == `camera`, fut/voxel_renderer.fut:5-12"""
    _fields_ = [(n, ctypes.c_float) for n in "x y height angle horizon distance fov".split()] + [
        ("sky_color", ctypes.c_uint32)]


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in "z0 delta invz_param1 invz_param2".split()] + [
        (n, ctypes.c_int32) for n in "filter sentinel f2i_mode flags".split()]


class FsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fsb error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile the library in-tree (nvcc -gencode arch=compute_100a,code=sm_100a; see csrc/Makefile)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None

# name -> (restype, argtypes); every symbol include/futspace_b200.h declares.
_vp, _ci, _cf, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_P = ctypes.POINTER
SYMBOLS = {
    "fsb_context_new": (_ci, [_ci, _P(_vp)]),
    "fsb_context_free": (None, [_vp]),
    "fsb_context_get_error": (ctypes.c_char_p, [_vp]),
    "fsb_context_sync": (_ci, [_vp]),
    "fsb_context_stream": (_vp, [_vp]),
    "fsb_context_launch_count": (ctypes.c_int64, [_vp]),
    "fsb_context_device_name": (_ci, [_vp, ctypes.c_char_p, _sz]),
    "fsb_context_set_profiling": (_ci, [_vp, _ci]),
    "fsb_context_get_profile": (_ci, [_vp, _P(ctypes.c_double), _P(ctypes.c_int64)]),
    "fsb_context_get_counters": (_ci, [_vp, _P(ctypes.c_uint64), _P(ctypes.c_uint64)]),
    "fsb_context_paint_trips": (ctypes.c_uint64, [_vp]),
    "fsb_params_default": (None, [_P(Params)]),
    "fsb_params_tests_variant": (None, [_P(Params)]),
    "fsb_get_zs": (_ci, [_cf, _cf, _cf, _vp, _ci]),
    "fsb_map_new": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _P(_vp)]),
    "fsb_map_new_split": (_ci, [_vp, _vp, _ci, _ci, _vp, _ci, _ci, _ci, _P(_vp)]),
    "fsb_map_free": (_ci, [_vp, _vp]),
    "fsb_map_is_packed": (_ci, [_vp]),
    "fsb_map_bake_shadows": (_ci, [_vp, _vp, _P(_cf), _ci, _ci, _vp]),
    "fsb_sun_vector": (None, [_cf, _cf, _P(_cf)]),
    "fsb_render": (_ci, [_vp, _P(Camera), _P(Params), _vp, _ci, _ci, _vp]),
    "fsb_render_device": (_ci, [_vp, _P(Camera), _P(Params), _vp, _ci, _ci, _vp, ctypes.c_int64]),
    "fsb_render_batch_device": (_ci, [_vp, _P(Camera), _ci, _P(Params), _vp, _ci, _ci, _vp]),
    "fsb_render_batch": (_ci, [_vp, _P(Camera), _ci, _P(Params), _vp, _ci, _ci, _vp]),
    "fsb_render_columns_device": (_ci, [_vp, _P(Camera), _P(Params), _vp, _ci, _ci, _ci, _ci, _vp, ctypes.c_int64]),
    "fsb_render_columns": (_ci, [_vp, _P(Camera), _P(Params), _vp, _ci, _ci, _ci, _ci, _vp, ctypes.c_int64]),
    "fsb_effect_interpolate_device": (_ci, [_vp, _ci, _vp, _ci, _ci, _vp]),
    "fsb_effect_interpolate2_device": (_ci, [_vp, _vp, _ci, _ci, _vp]),
    "fsb_device_malloc": (_ci, [_vp, _sz, _P(_vp)]),
    "fsb_device_free": (_ci, [_vp, _vp]),
    "fsb_host_malloc": (_ci, [_vp, _sz, _P(_vp)]),
    "fsb_host_free": (_ci, [_vp, _vp]),
    "fsb_host_register": (_ci, [_vp, _vp, _sz]),
    "fsb_host_unregister": (_ci, [_vp, _vp]),
    "fsb_host_is_registered": (_ci, [_vp, _vp]),
    "fsb_copy_to_host": (_ci, [_vp, _vp, _vp, _sz]),
    "fsb_copy_to_device": (_ci, [_vp, _vp, _vp, _sz]),
    "fsb_ipc_export": (_ci, [_vp, _vp, _vp]),
    "fsb_ipc_import": (_ci, [_vp, _vp, _P(_vp)]),
    "fsb_ipc_close": (_ci, [_vp, _vp]),
    "fsb_terrain_fbm": (_ci, [_ci, ctypes.c_uint64, _vp, _vp]),
    "fsb_selftest_sqrt": (_ci, [_vp, ctypes.c_uint32, ctypes.c_uint32, _P(ctypes.c_uint64)]),
    "fsb_bench_l2_stream": (_ci, [_vp, _sz, _ci, _P(ctypes.c_double)]),
    "fsb_bench_l2_gather": (_ci, [_vp, _sz, _ci, _P(ctypes.c_double)]),
}


def lib():
    """Load libfutspace_b200.so.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "futspace_b200 has no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (rt, at) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = rt, at
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().fsb_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def tests_variant_params(**kw):
    p = Params()
    lib().fsb_params_tests_variant(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def get_zs(delta, distance, z0, cap=1 << 22):
    n = lib().fsb_get_zs(delta, distance, z0, None, 0)
    if n < 0:
        raise ValueError("z-series undefined")
    buf = np.zeros(max(n, 1), np.float32)
    lib().fsb_get_zs(delta, distance, z0, buf.ctypes.data, min(n, cap))
    return buf[:n]


def sun_vector(sun_height, sun_ang):
    v = (_cf * 3)()
    lib().fsb_sun_vector(sun_height, sun_ang, v)
    return [v[0], v[1], v[2]]


def terrain_fbm(m, seed=0x5EED5EED):
    color = np.empty((m, m), np.uint32)
    height = np.empty((m, m), np.int32)
    rc = lib().fsb_terrain_fbm(m, seed, color.ctypes.data, height.ctypes.data)
    if rc:
        raise FsbError(rc, "fsb_terrain_fbm(m=%d)" % m)
    return color, height


class Map:
    def __init__(self, ctx, handle, q, r):
        self.ctx, self.handle, self.q, self.r = ctx, handle, q, r

    @property
    def packed(self):
        return bool(lib().fsb_map_is_packed(self.handle))

    def free(self):
        if self.handle:
            self.ctx._check(lib().fsb_map_free(self.ctx.handle, self.handle))
            self.handle = None


class Context:
    """One GPU, one stream (futhark_context analogue)."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib().fsb_context_new(device, ctypes.byref(h))
        if rc:
            raise FsbError(rc, "fsb_context_new(device=%d) failed%s" % (
                device, ": no sm_100 GPU (no CPU fallback exists)" if rc == ERR_NO_DEVICE else ""))
        self.handle = h

    def close(self):
        if self.handle:
            lib().fsb_context_free(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise FsbError(rc, lib().fsb_context_get_error(self.handle).decode())

    def error(self):
        return lib().fsb_context_get_error(self.handle).decode()

    def sync(self):
        self._check(lib().fsb_context_sync(self.handle))

    @property
    def stream(self):
        return lib().fsb_context_stream(self.handle)

    @property
    def launch_count(self):
        return lib().fsb_context_launch_count(self.handle)

    @property
    def device_name(self):
        b = ctypes.create_string_buffer(128)
        lib().fsb_context_device_name(self.handle, b, 128)
        return b.value.decode()

    def set_profiling(self, enable):
        self._check(lib().fsb_context_set_profiling(self.handle, 1 if enable else 0))

    def get_profile(self):
        """-> {"setup"|"march"|"colour"|"expand": (total_ms, launch groups)} since the last call."""
        ms = (ctypes.c_double * 4)()
        n = (ctypes.c_int64 * 4)()
        self._check(lib().fsb_context_get_profile(self.handle, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(("setup", "march", "colour", "expand"))}

    def get_counters(self):
        """-> (chunks_evaluated, records) since the last call (profiling must be on)."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(lib().fsb_context_get_counters(self.handle, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def paint_trips(self):
        """warp-wide colour trips of the paint kernel in the interval the last get_counters() closed"""
        return int(lib().fsb_context_paint_trips(self.handle))

    def upload_map(self, color, height, mask_heights=True):
        color = np.ascontiguousarray(color, dtype=np.uint32)
        height = np.ascontiguousarray(height, dtype=np.int32)
        if color.ndim != 2 or color.shape != height.shape:
            raise ValueError("colour and height maps must be 2-D and the same shape")
        h = _vp()
        self._check(lib().fsb_map_new(self.handle, color.ctypes.data, height.ctypes.data, color.shape[0],
                                      color.shape[1], 1 if mask_heights else 0, ctypes.byref(h)))
        return Map(self, h, color.shape[0], color.shape[1])

    def upload_map_split(self, color, height, mask_heights=True):
        """fsb_map_new_split: colour and height maps of different sizes (each wraps by its own size)."""
        color = np.ascontiguousarray(color, dtype=np.uint32)
        height = np.ascontiguousarray(height, dtype=np.int32)
        h = _vp()
        self._check(lib().fsb_map_new_split(self.handle, color.ctypes.data, color.shape[0], color.shape[1],
                                            height.ctypes.data, height.shape[0], height.shape[1],
                                            1 if mask_heights else 0, ctypes.byref(h)))
        return Map(self, h, height.shape[0], height.shape[1])

    def bake_shadows(self, mp, sun, out_q=None, out_r=None):
        """fsb_map_bake_shadows -> shadowed colour map [out_q][out_r] (numpy u32)."""
        out_q, out_r = out_q or mp.q, out_r or mp.r
        out = np.empty((out_q, out_r), np.uint32)
        s = (_cf * 3)(*sun)
        self._check(lib().fsb_map_bake_shadows(self.handle, mp.handle, s, out_q, out_r, out.ctypes.data))
        return out

    def render(self, cam, prm, mp, h, w, out=None):
        """fsb_render: host frame out (blocking)."""
        if out is None:
            out = np.empty((h, w), np.uint32)
        self._check(lib().fsb_render(self.handle, ctypes.byref(cam), ctypes.byref(prm), mp.handle, h, w,
                                     out.ctypes.data))
        return out

    def render_batch(self, cams, prm, mp, h, w, out=None):
        arr = cams if isinstance(cams, ctypes.Array) else (Camera * len(cams))(*cams)
        if out is None:
            out = np.empty((len(arr), h, w), np.uint32)
        ptr = out if isinstance(out, int) else out.ctypes.data
        self._check(lib().fsb_render_batch(self.handle, arr, len(arr), ctypes.byref(prm), mp.handle, h, w, ptr))
        return out

    def render_device(self, cam, prm, mp, h, w, out_dev, row_stride=0):
        self._check(lib().fsb_render_device(self.handle, ctypes.byref(cam), ctypes.byref(prm), mp.handle, h, w,
                                            out_dev, row_stride))

    def render_batch_device(self, cams, prm, mp, h, w, out_dev):
        arr = cams if isinstance(cams, ctypes.Array) else (Camera * len(cams))(*cams)
        self._check(lib().fsb_render_batch_device(self.handle, arr, len(arr), ctypes.byref(prm), mp.handle, h, w,
                                                  out_dev))

    def render_columns_device(self, cam, prm, mp, h, w, col_begin, col_end, out_dev, row_stride=0):
        self._check(lib().fsb_render_columns_device(self.handle, ctypes.byref(cam), ctypes.byref(prm), mp.handle,
                                                    h, w, col_begin, col_end, out_dev, row_stride))

    def render_columns(self, cam, prm, mp, h, w, col_begin, col_end, out_host, row_stride=0):
        """fsb_render_columns: slab -> host memory at `out_host` (address of pixel (0, col_begin)), blocking."""
        self._check(lib().fsb_render_columns(self.handle, ctypes.byref(cam), ctypes.byref(prm), mp.handle, h, w,
                                             col_begin, col_end, out_host, row_stride))

    def host_register_ptr(self, ptr, nbytes):
        self._check(lib().fsb_host_register(self.handle, ptr, nbytes))

    def host_unregister_ptr(self, ptr):
        self._check(lib().fsb_host_unregister(self.handle, ptr))

    def effect_interpolate(self, frame, pd=None):
        """fut/effects.fut post-passes on a host frame: interpolate2 (pd None) or interpolate pd."""
        frame = np.ascontiguousarray(frame, np.uint32)
        h, w = frame.shape
        a, b = self.device_malloc(frame.nbytes), self.device_malloc(frame.nbytes)
        try:
            self._check(lib().fsb_copy_to_device(self.handle, a, frame.ctypes.data, frame.nbytes))
            if pd is None:
                self._check(lib().fsb_effect_interpolate2_device(self.handle, a, h, w, b))
            else:
                self._check(lib().fsb_effect_interpolate_device(self.handle, pd, a, h, w, b))
            return self.download(b, (h, w))
        finally:
            self.device_free(a)
            self.device_free(b)

    def device_malloc(self, nbytes):
        p = _vp()
        self._check(lib().fsb_device_malloc(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._check(lib().fsb_device_free(self.handle, ptr))

    def host_malloc(self, nbytes):
        p = _vp()
        self._check(lib().fsb_host_malloc(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def host_free(self, ptr):
        self._check(lib().fsb_host_free(self.handle, ptr))

    def host_register(self, array):
        """fsb_host_register on a numpy array the caller keeps alive (unregister before dropping it)."""
        self._check(lib().fsb_host_register(self.handle, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        self._check(lib().fsb_host_unregister(self.handle, array.ctypes.data))

    def copy_to_host(self, dst, src_dev, nbytes):
        self._check(lib().fsb_copy_to_host(self.handle, dst, src_dev, nbytes))

    def download(self, src_dev, shape):
        out = np.empty(shape, np.uint32)
        self.copy_to_host(out.ctypes.data, src_dev, out.nbytes)
        self.sync()
        return out

    def selftest_sqrt(self, lo_bits, hi_bits):
        v = ctypes.c_uint64()
        self._check(lib().fsb_selftest_sqrt(self.handle, lo_bits, hi_bits, ctypes.byref(v)))
        return v.value

    def ipc_export(self, dev_ptr):
        h = ctypes.create_string_buffer(64)
        self._check(lib().fsb_ipc_export(self.handle, dev_ptr, h))
        return h.raw

    def ipc_import(self, handle_bytes):
        p = _vp()
        self._check(lib().fsb_ipc_import(self.handle, ctypes.create_string_buffer(handle_bytes, 64), ctypes.byref(p)))
        return p.value

    def ipc_close(self, dev_ptr):
        self._check(lib().fsb_ipc_close(self.handle, dev_ptr))

    def l2_stream_gbs(self, nbytes=48 << 20, iters=20):
        v = ctypes.c_double()
        self._check(lib().fsb_bench_l2_stream(self.handle, nbytes, iters, ctypes.byref(v)))
        return v.value

    def l2_gather_gsectors(self, nbytes=64 << 20, iters=20):
        v = ctypes.c_double()
        self._check(lib().fsb_bench_l2_gather(self.handle, nbytes, iters, ctypes.byref(v)))
        return v.value
