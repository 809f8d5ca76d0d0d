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
for flags in (0, F.FLAG_NO_TEXTURE, F.FLAG_FORCE_GENERIC, F.FLAG_NO_CULL, F.FLAG_MARCH_Z):
    for filt in (0, 1):
        prm = F.default_params(filter=filt, flags=flags)
        ctx.render(cams[0], prm, mp, 130, 170)
        ctx.render_batch(cams, prm, mp, 97, 65)
# round 2: column-parallel march + colour pass (forced for small groups), colour slices, single-frame march on/off
# + the paint kernel (colour pass and expand as one kernel): whole columns, segments of bands, and the two launches it replaced
for env in ({"FSB_COLS_MIN_WARPS": "0", "FSB_PAINT_SEG": "0"}, {"FSB_COLS_MIN_WARPS": "0", "FSB_PAINT_SEG": "2"},
            {"FSB_COLS_MIN_WARPS": "0"},
            {"FSB_COLS_MIN_WARPS": "0", "FSB_PAINT": "0", "FSB_COLOUR_SLICE": "0"},
            {"FSB_COLS_MIN_WARPS": "0", "FSB_PAINT": "0", "FSB_COLOUR_SLICE": "7"},
            {"FSB_FRAME_MAX_COLS": "0"}, {"FSB_FRAME_MAX_COLS": "100000000"}):
    os.environ.update(env)
    for filt in (0, 1):
        for flags in (0, F.FLAG_NO_CULL, F.FLAG_SMOOTHING):
            prm = F.default_params(filter=filt, flags=flags)
            ctx.render(cams[0], prm, mp, 130, 170)
            ctx.render_batch(cams, prm, mp, 97, 65)
    ctx.render(cams[2], F.tests_variant_params(), mp, 300, 40)
    for k in env:
        del os.environ[k]
dev = ctx.device_malloc(130 * 176 * 4)      # column slabs (aligned and unaligned destinations: TMA and per-lane stores)
for c0, c1 in ((0, 64), (64, 101), (101, 170)):
    ctx.render_columns_device(cams[0], F.default_params(), mp, 130, 170, c0, c1, dev + 4 * c0, 176)
ctx.device_free(dev)
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
