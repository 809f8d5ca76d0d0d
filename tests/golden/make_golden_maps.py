#!/usr/bin/env python3
"""Commits three more of the reference's 29 map pairs as fixtures (run in the build container only).

SURVEY.md section 4 asks for bit-comparison on the real map pairs at the tests/futspace.fut camera; /root/reference does
not exist on the GPU box, so a few pairs travel as palette-indexed planes (the converted maps hold <= 256 colours):
  c{n}w_d{n}_pal.npz : idx [1024][1024] u8, pal [<=256] u32 (0x00RRGGBB, tools/png2data.py:45-48), height [1024][1024] u8
                       (grey level of the D map, tools/png2data-grey.py:22-24)
"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data/converted-maps"

for n in (7, 13, 29):
    c = np.array(Image.open(os.path.join(REF, "C%dW.png" % n)).convert("RGB"))
    d = np.array(Image.open(os.path.join(REF, "D%d.png" % n)).convert("RGB"))
    assert (d[..., 0] == d[..., 1]).all() and (d[..., 1] == d[..., 2]).all()
    rgb = (c[..., 0].astype(np.uint32) << 16) | (c[..., 1].astype(np.uint32) << 8) | c[..., 2]
    pal = np.unique(rgb)
    assert len(pal) <= 256
    idx = np.searchsorted(pal, rgb).astype(np.uint8)
    assert (pal[idx] == rgb).all()
    out = os.path.join(HERE, "c%dw_d%d_pal.npz" % (n, n))
    np.savez_compressed(out, idx=idx, pal=pal, height=d[..., 2])
    print(out, os.path.getsize(out) >> 10, "KB", len(pal), "colours, max height", d[..., 2].max())
