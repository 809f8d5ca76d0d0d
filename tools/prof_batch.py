"""Batch workload for ncu: N poses of the 1080p camera path, a few launches."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "1080p"]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
ctx = F.Context(0)
col, hgt = F.terrain_fbm(m)
mp = ctx.upload_map(col, hgt)
prm = F.default_params()
cams = bench.camera_path(F, hgt, m, 512, 0, P, h, dist)
arr = (F.Camera * P)(*cams)
dev = ctx.device_malloc(P * h * w * 4)
for _ in range(reps):
    ctx.render_batch_device(arr, prm, mp, h, w, dev)
ctx.sync()
