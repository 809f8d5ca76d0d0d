"""ctypes view of include/libfutspace.h (the generated-API names) for the tests: drives the library the way
c/interactive.c and lys do."""
import ctypes

import numpy as np

import futspace_b200 as F

vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
i32, u32, i64 = ctypes.c_int32, ctypes.c_uint32, ctypes.c_int64
P = ctypes.POINTER

SDLK_UP, SDLK_DOWN, SDLK_LEFT, SDLK_RIGHT = 0x40000052, 0x40000051, 0x40000050, 0x4000004F


def lib():
    L = ctypes.CDLL(F.LIB_PATH)
    sig = {
        "futhark_context_config_new": (vp, []), "futhark_context_config_free": (None, [vp]),
        "futhark_context_config_set_device": (None, [vp, ctypes.c_char_p]),
        "futhark_context_new": (vp, [vp]), "futhark_context_free": (None, [vp]),
        "futhark_context_sync": (ci, [vp]), "futhark_context_get_error": (vp, [vp]),
        "futhark_new_u32_2d": (vp, [vp, vp, i64, i64]), "futhark_free_u32_2d": (ci, [vp, vp]),
        "futhark_values_u32_2d": (ci, [vp, vp, vp]), "futhark_shape_u32_2d": (P(i64), [vp, vp]),
        "futhark_new_i32_2d": (vp, [vp, vp, i64, i64]), "futhark_free_i32_2d": (ci, [vp, vp]),
        "futhark_free_opaque_state": (ci, [vp, vp]),
        "futhark_entry_init": (ci, [vp, P(vp), u32]),
        "futhark_entry_resize": (ci, [vp, P(vp), i32, i32, vp]),
        "futhark_entry_key": (ci, [vp, P(vp), i32, i32, vp]),
        "futhark_entry_mouse": (ci, [vp, P(vp), i32, i32, i32, vp]),
        "futhark_entry_wheel": (ci, [vp, P(vp), i32, i32, vp]),
        "futhark_entry_step": (ci, [vp, P(vp), cf, vp]),
        "futhark_entry_render": (ci, [vp, P(vp), vp]),
        "futhark_entry_text_content": (ci, [vp] + [P(cf)] * 9 + [vp]),
        "futhark_entry_update_map": (ci, [vp, P(vp), vp, vp, vp]),
    }
    for n, (rt, at) in sig.items():
        f = getattr(L, n)
        f.restype, f.argtypes = rt, at
    L._libc_free = ctypes.CDLL(None).free
    L._libc_free.argtypes = [vp]
    return L


class Session:
    """The call sequence of c/interactive.c: context, init, (update_map), key/step/render."""

    def __init__(self, device="0"):
        self.L = lib()
        self.cfg = self.L.futhark_context_config_new()
        self.L.futhark_context_config_set_device(self.cfg, device.encode())
        self.ctx = self.L.futhark_context_new(self.cfg)
        self.state = None

    def error(self):
        p = self.L.futhark_context_get_error(self.ctx)
        if not p:
            return None
        s = ctypes.string_at(p).decode()
        self.L._libc_free(p)
        return s

    def _swap(self, new):
        if self.state:
            self.L.futhark_free_opaque_state(self.ctx, self.state)
        self.state = new

    def _entry(self, fn, *args):
        out = vp()
        rc = fn(self.ctx, ctypes.byref(out), *args)
        if rc:
            raise RuntimeError(self.error())
        self._swap(out)

    def init(self, seed=1):
        self._entry(self.L.futhark_entry_init, seed)

    def resize(self, h, w):
        self._entry(self.L.futhark_entry_resize, h, w, self.state)

    def key(self, down, key):
        self._entry(self.L.futhark_entry_key, 0 if down else 1, key, self.state)

    def step(self, td=0.016):
        self._entry(self.L.futhark_entry_step, td, self.state)

    def mouse(self):
        self._entry(self.L.futhark_entry_mouse, 0, 1, 2, self.state)

    def update_map(self, color, height):
        color = np.ascontiguousarray(color, np.uint32)
        height = np.ascontiguousarray(height, np.int32)
        c = self.L.futhark_new_u32_2d(self.ctx, color.ctypes.data, color.shape[0], color.shape[1])   # c/interactive.c:50
        h = self.L.futhark_new_i32_2d(self.ctx, height.ctypes.data, height.shape[0], height.shape[1])  # :51
        try:
            self._entry(self.L.futhark_entry_update_map, c, h, self.state)                            # :54
        finally:
            self.L.futhark_free_u32_2d(self.ctx, c)                                                    # :55-56
            self.L.futhark_free_i32_2d(self.ctx, h)

    def text_content(self):
        v = [cf() for _ in range(9)]
        rc = self.L.futhark_entry_text_content(self.ctx, *[ctypes.byref(x) for x in v], self.state)
        assert rc == 0
        return [x.value for x in v]

    def render(self):
        out = vp()
        if self.L.futhark_entry_render(self.ctx, ctypes.byref(out), self.state):
            raise RuntimeError(self.error())
        shp = self.L.futhark_shape_u32_2d(self.ctx, out)
        frame = np.empty((shp[0], shp[1]), np.uint32)
        assert self.L.futhark_values_u32_2d(self.ctx, out, frame.ctypes.data) == 0
        assert self.L.futhark_context_sync(self.ctx) == 0
        self.L.futhark_free_u32_2d(self.ctx, out)
        return frame

    def close(self):
        self._swap(None)
        self.L.futhark_context_free(self.ctx)
        self.L.futhark_context_config_free(self.cfg)
