"""Per-kernel device time of single frames (set-up / march / expand) through the library's own event profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
import bench

ctx = F.Context(0)
for name in ("cfg1", "1080p", "4k"):
    wl = bench.WORKLOADS[name]
    m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
    col, hgt = F.terrain_fbm(m)
    mp = ctx.upload_map(col, hgt)
    cam = F.Camera(m / 2 + 0.37, m / 2 + 0.73, max(160.0, float(hgt[m // 2, m // 2]) + 20.0), 2.2, 0.3 * h, dist, 1.2, bench.SKY)
    dev = ctx.device_malloc(w * h * 4)
    for flags in (0, F.FLAG_NO_CULL):
        prm = F.default_params(flags=flags)
        for _ in range(5):
            ctx.render_device(cam, prm, mp, h, w, dev)
        ctx.set_profiling(True)
        ctx.render_device(cam, prm, mp, h, w, dev)
        ctx.get_profile()
        for _ in range(20):
            ctx.render_device(cam, prm, mp, h, w, dev)
        p = ctx.get_profile()
        ctx.set_profiling(False)
        print(name, "cull off" if flags else "default ", {k: "%.1f us" % (1e3 * v[0] / v[1]) for k, v in p.items()})
    ctx.device_free(dev)
    mp.free()
