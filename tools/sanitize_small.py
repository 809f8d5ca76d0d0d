"""Small renders covering every kernel variant, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import futspace_b200 as F
SKY = 0xFF9090E0
ctx = F.Context(0)
col, hgt = F.terrain_fbm(256)
mp = ctx.upload_map(col, hgt)
mp2 = ctx.upload_map(col[:200, :133].copy() ^ 0x01000000 * (np.arange(133) & 1).astype(np.uint32), hgt[:200, :133].copy() * 3, mask_heights=False)
cams = [F.Camera(100.3 + 5 * i, 77.7, 200, 2.2 + 0.2 * i, 60, 150 + 10 * i, 1.2, SKY) for i in range(5)]
for flags in (0, F.FLAG_NO_TEXTURE, F.FLAG_FORCE_GENERIC, F.FLAG_NO_CULL):
    for filt in (0, 1):
        prm = F.default_params(filter=filt, flags=flags)
        ctx.render(cams[0], prm, mp, 130, 170)
        ctx.render_batch(cams, prm, mp, 97, 65)
for f2i in (0, 1, 2):
    ctx.render(cams[1], F.default_params(f2i_mode=f2i), mp2, 64, 48)
ctx.render(cams[2], F.tests_variant_params(), mp, 300, 40)
for flags in (F.FLAG_SMOOTHING, F.FLAG_SMOOTHING | F.FLAG_FORCE_GENERIC, F.FLAG_SMOOTHING | F.FLAG_NO_TEXTURE):
    ctx.render(cams[3], F.default_params(flags=flags), mp, 130, 170)
    ctx.render_batch(cams, F.default_params(filter=0, flags=flags), mp, 33, 65)
frame = ctx.render(cams[0], F.default_params(), mp, 64, 96)
ctx.effect_interpolate(frame, 2)
ctx.effect_interpolate(frame)
ctx.bake_shadows(mp, F.sun_vector(1.2, 0.4), 64, 64)
print("sanitize workload done, launches:", ctx.launch_count)
