"""fsb_render into pageable vs registered vs fsb_host_malloc'ed memory: time per frame (host wall clock, blocking call)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import futspace_b200 as F
ctx = F.Context(0)
col, hgt = F.terrain_fbm(2048)
mp = ctx.upload_map(col, hgt)
prm = F.default_params()
for (h, w, dist) in ((1024, 1024, 800.0), (1080, 1920, 2000.0), (2160, 3840, 4000.0)):
    cam = F.Camera(1024.37, 1024.73, 200, 2.2, 0.3 * h, dist, 1.2, 0xFF9090E0)
    buf = np.zeros((h, w), np.uint32)
    def run(n=30):
        ctx.render(cam, prm, mp, h, w, out=buf)
        t = time.perf_counter()
        for _ in range(n):
            ctx.render(cam, prm, mp, h, w, out=buf)
        return (time.perf_counter() - t) / n * 1e6
    a = run()
    ctx.host_register(buf)
    b = run()
    ctx.host_unregister(buf)
    print("%dx%d: fsb_render into pageable memory %.0f us/frame, into the same buffer registered %.0f us/frame (%.1f MB frame)" % (w, h, a, b, h * w * 4 / 1e6))
