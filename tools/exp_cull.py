import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F, bench
wl = bench.WORKLOADS["1080p"]; m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
ctx = F.Context(0); col, hgt = F.terrain_fbm(m); mp = ctx.upload_map(col, hgt)
dev = ctx.device_malloc(h*w*4)
ctx.set_profiling(True)
for pi in (0, 61, 122, 183, 244, 305):
    cam = bench.camera_path(F, hgt, m, 512, pi, 1, h, dist)[0]
    out = []
    for flags in (0, F.FLAG_NO_CULL):
        ctx.render_device(cam, F.default_params(flags=flags), mp, h, w, dev); ctx.sync()
        out.append(ctx.get_counters()[0])
    print(pi, "cam_h %.1f" % cam.height, "evaluated frac", out[0]/out[1])
