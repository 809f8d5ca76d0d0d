"""CPU estimate: how many (group of 32 columns, chunk of 32 samples) blocks of the column-parallel march could a
local-maximum bound skip?  Nearest-neighbour heights (statistics only, not parity)."""
import sys, os, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import oracle_lib as O
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "1080p"]
m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
col, hgt = O.terrain_fbm(m)
hgt = (hgt & 0xFF).astype(np.float32)
hmax = hgt.max()
prm = O.default_params()
zs = np.array(O.get_zs(prm.delta, dist, prm.z0), np.float64)
nz = len(zs)
nch = (nz + 31) // 32
# max-mip pyramid (block maxima), level L = 2^L texels
pyr = [hgt]
while pyr[-1].shape[0] > 1:
    a = pyr[-1]
    pyr.append(np.maximum(np.maximum(a[0::2, 0::2], a[1::2, 0::2]), np.maximum(a[0::2, 1::2], a[1::2, 1::2])))
tot_blocks = ev_now = skip_exact = 0
skip_mip = {}
for pi in range(0, 512, 64):
    cam = bench.camera_path(O, hgt, m, 512, pi, 1, h, dist)[0]
    s, c = math.sin(cam.angle), math.cos(cam.angle)
    fov = cam.fov
    lx, ly = (-c - s * fov) * zs + cam.x, (s - c * fov) * zs + cam.y
    rx, ry = (c - s * fov) * zs + cam.x, (-s - c * fov) * zs + cam.y
    j = np.arange(w)[None, :]
    X = lx[:, None] + j * ((rx - lx) / w)[:, None]
    Y = ly[:, None] + j * ((ry - ly) / w)[:, None]
    xi, yi = np.floor(X).astype(np.int64) % m, np.floor(Y).astype(np.int64) % m
    H = hgt[yi, xi]
    # bilinear upper bound: max of the 2x2 footprint
    H = np.maximum(np.maximum(H, hgt[yi, (xi + 1) % m]), np.maximum(hgt[(yi + 1) % m, xi], hgt[(yi + 1) % m, (xi + 1) % m]))
    with np.errstate(divide="ignore", invalid="ignore"):
        iz = (1.0 / zs) * (w // 2)
        rows = np.clip(np.nan_to_num((cam.height - H) * iz[:, None] + cam.horizon, nan=0, posinf=1e9, neginf=0), 0, 1e9)
    rows = np.floor(rows)
    run = np.minimum.accumulate(np.vstack([np.full((1, w), h), rows]), axis=0)  # ybuf entering step k = run[k]
    # blocks
    ng = w // 32
    pad = nch * 32 - nz
    Hp = np.vstack([H, np.repeat(H[-1:], pad, 0)]) if pad else H
    izp = np.concatenate([iz, np.repeat(iz[-1:], pad)])
    ybuf_in = run[:-1]
    ybuf_in = np.vstack([ybuf_in, np.repeat(run[-1:], pad, 0)]) if pad else ybuf_in
    Hb = Hp.reshape(nch, 32, ng, 32)
    ymax_in = ybuf_in.reshape(nch, 32, ng, 32)[:, 0].max(axis=2)         # max over the group's lanes of ybuf at chunk entry
    iz_first, iz_last = izp.reshape(nch, 32)[:, 0], izp.reshape(nch, 32)[:, -1]
    def bound_row(hm):  # lowest row any sample of the block can have, given heights <= hm
        d = cam.height - (hm + 0.5)
        izsel = np.where(d >= 0, iz_last[:, None], iz_first[:, None])
        with np.errstate(invalid="ignore"):
            return np.maximum(0, np.floor(np.nan_to_num(d * izsel + cam.horizon, nan=0, posinf=1e9, neginf=-1e9)))
    # current rule (global maximum); "finished for good" also needs ybuf <= bound for the rest, approximated per block
    now_skip = bound_row(np.full((nch, ng), hmax)) >= ymax_in
    # after a block is skipped by the monotone global bound the column stays finished: emulate with cumulative OR for camera below
    ev = ~now_skip
    exact_skip = bound_row(Hb.max(axis=(1, 3))) >= ymax_in
    tot_blocks += nch * ng
    ev_now += ev.sum()
    skip_exact += (ev & exact_skip).sum()
    # mip bound: bbox of the block's sample positions, level with block size >= bbox/2 -> at most 3x3 texels of that level
    Xb = np.vstack([X, np.repeat(X[-1:], pad, 0)]).reshape(nch, 32, ng, 32) if pad else X.reshape(nch, 32, ng, 32)
    Yb = np.vstack([Y, np.repeat(Y[-1:], pad, 0)]).reshape(nch, 32, ng, 32) if pad else Y.reshape(nch, 32, ng, 32)
    for taps in (1, 4, 8):   # taps x taps sub-blocks per block, each bounded by its own bbox at its own mip level
        hm = np.zeros((nch, ng), np.float32)
        for ci in range(nch):
            for g in range(ng):
                if not ev[ci, g]:
                    continue
                best = 0.0
                st = 32 // taps
                for a0 in range(0, 32, st):
                    for b0 in range(0, 32, st):
                        xs, ys = Xb[ci, a0:a0 + st, g, b0:b0 + st], Yb[ci, a0:a0 + st, g, b0:b0 + st]
                        x0, x1, y0, y1 = math.floor(xs.min()), math.floor(xs.max()) + 1, math.floor(ys.min()), math.floor(ys.max()) + 1
                        ext = max(x1 - x0, y1 - y0) + 1
                        L = max(0, math.ceil(math.log2(ext)))
                        L = min(L, len(pyr) - 1)
                        P = pyr[L]
                        n = P.shape[0]
                        bx0, bx1, by0, by1 = x0 >> L, x1 >> L, y0 >> L, y1 >> L
                        for by in range(by0, by1 + 1):
                            for bx in range(bx0, bx1 + 1):
                                best = max(best, P[by % n, bx % n])
                hm[ci, g] = best
        sk = (ev & (bound_row(hm) >= ymax_in)).sum()
        skip_mip[taps] = skip_mip.get(taps, 0) + sk
    print(pi, "blocks", nch * ng, "evaluated now", ev.sum(), "exact-local skip", (ev & exact_skip).sum(), {k: int(v) for k, v in skip_mip.items()}, flush=True)
print("total blocks", tot_blocks, "evaluated now %.3f" % (ev_now / tot_blocks), "of those skippable with exact local max %.3f" % (skip_exact / ev_now),
      {k: "%.3f" % (v / ev_now) for k, v in skip_mip.items()})
