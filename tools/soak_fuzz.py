"""Randomised parity soak: GPU (through the C-ABI) against the oracle on many small random configurations.
usage: python tools/soak_fuzz.py [seconds] [seed]      (prints every mismatch with the iteration number to reproduce it)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import futspace_b200 as F
import oracle_lib as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tall = len(sys.argv) > 3 and sys.argv[3] == "tall"   # tall narrow frames, long depth series: many bands, long columns
rng = np.random.default_rng(seed)
ctx = F.Context(0)
base_col, base_hgt = F.terrain_fbm(512)
t0 = time.time()
it = bad = 0
kinds = {}
while time.time() - t0 < budget:
    it += 1
    # ---- a map ----
    q, r = int(rng.choice([1, 2, 3, 17, 64, 100, 128, 256, 333, 512])), int(rng.choice([1, 2, 5, 31, 64, 128, 200, 256, 512]))
    kind = int(rng.integers(0, 5))
    if kind == 0:      # terrain crop, alpha 0xFF
        col, hgt = base_col[:q, :r].copy(), base_hgt[:q, :r].copy()
    elif kind == 1:    # noise with transparent (colour 0) texels, alpha 0xFF elsewhere
        col = (rng.integers(0, 1 << 24, (q, r)).astype(np.uint32) | np.uint32(0xFF000000))
        col[rng.random((q, r)) < 0.2] = 0
        hgt = rng.integers(0, 256, (q, r)).astype(np.int32)
    elif kind == 2:    # uniform alpha 0 (tools/png2data.py maps) or an odd uniform alpha
        a = int(rng.choice([0, 0x80, 0x01, 0xFE]))
        col = rng.integers(0, 1 << 24, (q, r)).astype(np.uint32) | np.uint32(a << 24)
        hgt = (base_hgt[:q, :r] // 2).astype(np.int32)
    elif kind == 3:    # unpackable: heights beyond a byte, negative heights
        col = base_col[:q, :r].copy()
        hgt = (base_hgt[:q, :r].astype(np.int32) * 5 - 300)
    else:              # few colours, steps (lots of equal rows: ties in occlude)
        col = (np.uint32(0xFF000000) | (rng.integers(0, 4, (q, r)).astype(np.uint32) * np.uint32(0x404040)))
        hgt = (rng.integers(0, 6, (q, r)) * 40).astype(np.int32)
    mask = kind != 3
    mp = ctx.upload_map(col, hgt, mask_heights=mask)
    hm = (hgt & 0xFF) if mask else hgt
    for _ in range(6):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        if tall:
            h, w = int(rng.integers(200, 1300)), int(rng.integers(1, 48))
        cam = F.Camera(float(rng.uniform(-600, 600)), float(rng.uniform(-600, 600)), float(rng.uniform(-50, 500)),
                       float(rng.uniform(-7, 7)), float(rng.uniform(-50, h + 50)), float(rng.uniform(0.01, 900)),
                       float(rng.uniform(0.2, 2.5)), int(rng.integers(0, 1 << 32)))
        if tall:
            cam.distance = float(rng.uniform(500, 4500))
            cam.horizon = float(rng.uniform(0, h))
        if rng.random() < 0.15:
            cam.x, cam.y = float(int(cam.x)), float(int(cam.y))
        if rng.random() < 0.1:
            cam.x, cam.y = float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))
        u = rng.random()
        if u < 0.05:       # far outside the fast paths' coordinate range: generic kernel
            cam.x, cam.y = float(rng.uniform(-1e7, 1e7)), float(rng.uniform(-1e7, 1e7))
        elif u < 0.10:     # extreme heights / horizons
            cam.height, cam.horizon = float(rng.uniform(-1e4, 1e4)), float(rng.uniform(-1e4, 1e4))
        elif u < 0.15:     # degenerate view: zero / negative field of view, tiny distance (n_z = 0 or 1), huge angle
            cam.fov = float(rng.choice([0.0, -1.2, 1e-6]))
            cam.distance = float(rng.choice([1e-5, 0.0011, 3.0, cam.distance]))
            cam.angle = float(rng.uniform(-1e4, 1e4))
        elif u < 0.20:     # sentinel collisions: the sky colour is a colour of the map
            cam.sky_color = int(col[int(rng.integers(0, q)), int(rng.integers(0, r))])
        prm = F.default_params() if rng.random() < 0.7 else F.tests_variant_params()
        prm.filter = int(rng.integers(0, 2))
        prm.sentinel = int(rng.integers(0, 2))
        prm.f2i_mode = int(rng.choice([0, 0, 0, 1, 2]))
        prm.flags = int(rng.choice([0, 0, 0, 1, 2, 4, 8, 9, 10, 12, 16, 20, 24]))
        if prm.flags & 8:
            prm.sentinel = 0
        if rng.random() < 0.2:
            prm.invz_param1, prm.invz_param2 = float(rng.uniform(-1.0, 3.0)), float(rng.uniform(1, 600))
        if rng.random() < 0.15:
            prm.z0, prm.delta = float(rng.uniform(-1.0, 3.0)), float(rng.uniform(0.0005, 0.05))
        # round 2: which march / colour-pass shape renders it (read per call by the library)
        for k in ("FSB_COLS_MIN_WARPS", "FSB_COLOUR_SLICE", "FSB_FRAME_MAX_COLS", "FSB_PAINT", "FSB_PAINT_SEG"):
            os.environ.pop(k, None)
        v = int(rng.integers(0, 6))
        if v == 0:
            os.environ["FSB_COLS_MIN_WARPS"] = "0"
            os.environ["FSB_PAINT"] = "0"
            os.environ["FSB_COLOUR_SLICE"] = str(int(rng.choice([0, 1, 5, 32])))
        elif v == 5 or v == 3:   # column-parallel march + paint kernel: whole columns, or segments of 1..40 bands
            os.environ["FSB_COLS_MIN_WARPS"] = "0"
            os.environ["FSB_PAINT_SEG"] = str(int(rng.choice([0, 0, 1, 2, 3, 7, 40])))
        elif v == 1:
            os.environ["FSB_FRAME_MAX_COLS"] = "0"
        elif v == 2:
            os.environ["FSB_FRAME_MAX_COLS"] = "100000000"
        mode = int(rng.integers(0, 4))   # which entry point produces the frame
        try:
            if mode == 0:
                got = ctx.render(cam, prm, mp, h, w)
            elif mode == 1:                # batch of three poses, the checked one in a random slot
                others = [F.Camera(cam.x + 3.5, cam.y - 2.25, cam.height + 5, cam.angle + 0.3, cam.horizon, cam.distance * 0.5 + 1,
                                   cam.fov, cam.sky_color) for _ in range(2)]
                slot = int(rng.integers(0, 3))
                cams3 = others[:slot] + [cam] + others[slot:]
                got = ctx.render_batch(cams3, prm, mp, h, w)[slot]
            else:                          # column slabs rendered separately into one padded device frame
                stride = w + int(rng.integers(0, 5))
                dev = ctx.device_malloc(h * stride * 4)
                cut = int(rng.integers(0, w + 1))
                for c0, c1 in ((0, cut), (cut, w)):
                    if c1 > c0:
                        ctx.render_columns_device(cam, prm, mp, h, w, c0, c1, dev + c0 * 4, stride)
                got = ctx.download(dev, (h, stride))[:, :w].copy()
                ctx.device_free(dev)
        except F.FsbError as e:
            kinds["error"] = kinds.get("error", 0) + 1
            continue
        op = O.Params(prm.z0, prm.delta, prm.invz_param1, prm.invz_param2, prm.filter, prm.sentinel, prm.f2i_mode,
                      1 if prm.flags & 8 else 0)
        oc = O.Camera(cam.x, cam.y, cam.height, cam.angle, cam.horizon, cam.distance, cam.fov, cam.sky_color)
        want = O.render(oc, op, col, hm, h, w)
        kinds[kind] = kinds.get(kind, 0) + 1
        if not np.array_equal(got, want):
            bad += 1
            print("MISMATCH it=%d kind=%d map=%dx%d frame=%dx%d flags=%d filter=%d sentinel=%d f2i=%d differing=%d cam=%s"
                  % (it, kind, q, r, h, w, prm.flags, prm.filter, prm.sentinel, prm.f2i_mode, int((got != want).sum()),
                     [cam.x, cam.y, cam.height, cam.angle, cam.horizon, cam.distance, cam.fov]), flush=True)
    mp.free()
print("soak: seed %d, %d maps, renders per kind %s, %d mismatches, %.0f s" % (seed, it, kinds, bad, time.time() - t0))
sys.exit(1 if bad else 0)
