"""Scratch timing of the render kernel (device-resident output) for development."""
import ctypes, math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import futspace_b200 as F

SKY = 0xFF9090E0
def run(ctx, m, h, w, dist, n=20, prm=None, label=""):
    col, hgt = F.terrain_fbm(m)
    mp = ctx.upload_map(col, hgt)
    prm = prm or F.default_params()
    cam = F.Camera(m/2+0.37, m/2+0.73, 200, 2.2, 0.3*h, dist, 1.2, SKY)
    dev = ctx.device_malloc(h*w*4)
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3): ctx.render_device(cam, prm, mp, h, w, dev)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): ctx.render_device(cam, prm, mp, h, w, dev)
    e1.record(st); e1.synchronize()
    ms = e0.elapsed_time(e1)/n
    nz = len(F.get_zs(prm.delta, dist, prm.z0))
    taps = 4 if prm.filter else 1
    balg = 4*taps*w*nz + 4*w*h
    print(f"{label} {w}x{h} map {m} dist {dist} nz {nz}: {ms*1e3:.1f} us/frame  {1e3/ms:.0f} fps  {w*h/ms/1e3:.0f} Mpix/s  B_alg {balg/1e6:.1f} MB -> {balg/ms/1e6:.0f} GB/s")
    ctx.device_free(dev); mp.free()
    return ms

if __name__ == "__main__":
    ctx = F.Context(0)
    print(ctx.device_name)
    print("l2_stream GB/s", ctx.l2_stream_gbs(), "l2_gather Gsect/s", ctx.l2_gather_gsectors())
    run(ctx, 1024, 768, 1024, 1000, label="cfg1")
    run(ctx, 2048, 1080, 1920, 2000, label="cfg2")
    run(ctx, 4096, 2160, 3840, 4000, label="cfg3")
    run(ctx, 4096, 2160, 3840, 4000, prm=F.tests_variant_params(), label="cfg3-testsvariant")
