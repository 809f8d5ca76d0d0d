"""The reference's interactive loop (lys: step, render, values, sync) through include/libfutspace.h at the default
1024x1024 / distance 800 settings on the C1W/D1 map pair -- frames per second a host application would see."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import futhark_shim as FS
z = np.load(os.path.join(ROOT, "tests", "golden", "c1w_d1.npz"))
rgb = (z["r"].astype(np.uint32) << 16) | (z["g"].astype(np.uint32) << 8) | z["b"].astype(np.uint32) | 0xFF000000
s = FS.Session(); s.init(); s.update_map(rgb, z["height"].astype(np.int32))
s.key(True, ord("w")); s.key(True, ord("a"))
for _ in range(20):
    s.step(); s.render()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    s.step(); s.render()
dt = time.perf_counter() - t0
print("shim loop 1024x1024 d800: %.0f frames/s (%.1f us per step+render+values+sync)" % (n / dt, 1e6 * dt / n))
s.close()
