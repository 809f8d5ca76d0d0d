"""One workload, a few launches: the command ncu wraps (development + profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
SKY = 0xFF9090E0
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
m, h, w, dist = {"cfg1": (1024, 768, 1024, 1000), "cfg2": (2048, 1080, 1920, 2000), "cfg3": (4096, 2160, 3840, 4000)}[cfg]
variant = sys.argv[2] if len(sys.argv) > 2 else "live"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = F.Context(0)
col, hgt = F.terrain_fbm(m)
mp = ctx.upload_map(col, hgt)
prm = F.default_params() if variant == "live" else F.tests_variant_params()
cam = F.Camera(m/2+0.37, m/2+0.73, 200, 2.2, 0.3*h, dist, 1.2, SKY)
dev = ctx.device_malloc(h*w*4)
for _ in range(n):
    ctx.render_device(cam, prm, mp, h, w, dev)
ctx.sync()
