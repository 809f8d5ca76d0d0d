#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/ (run in the build container only).

Inputs : the reference's own map pair data/converted-maps/C1W.png + D1.png (read from
         /root/reference, which does not exist on the GPU box -- hence the committed copy).
Outputs: c1w_d1.npz      colour (0x00RRGGBB, the tools/png2data.py:45-48 convention) and height
                         (blue byte of the grey D map == tools/png2data-grey.py:22-24) as uint8 planes
         golden_frames.npz  frames produced by the ORACLE (oracle/fs_oracle.c) at fixed cameras.

The reference cannot be executed here (no futhark), so these frames are NOT reference outputs:
they pin the oracle against regressions ("parity unpinned").  The only numbers that
come from the reference itself are the z-series known answers checked in tests/test_oracle.py.
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REF = "/root/reference/data/converted-maps"


def main():
    c = np.array(Image.open(os.path.join(REF, "C1W.png")).convert("RGB"))
    d = np.array(Image.open(os.path.join(REF, "D1.png")).convert("RGB"))
    assert (d[..., 0] == d[..., 1]).all() and (d[..., 1] == d[..., 2]).all()
    np.savez_compressed(os.path.join(HERE, "c1w_d1.npz"), r=c[..., 0], g=c[..., 1], b=c[..., 2], height=d[..., 2])

    rgb = (c[..., 0].astype(np.uint32) << 16) | (c[..., 1].astype(np.uint32) << 8) | c[..., 2]
    hgt = d[..., 2].astype(np.int32)
    frames = {}
    # tests/futspace.fut:126-144 : fixed camera, 400x800, nearest, sky sentinel, no alpha in colours
    cam = O.Camera(512, 800, 78, 0, 100, 800, 1, 0xFF9090E0)
    frames["tests_variant_400x800"] = O.render(cam, O.tests_variant_params(), rgb, hgt, 400, 800)
    # live renderer at the init camera (fut/interactive.fut:29-36), loader alpha 0xFF
    # (c/freeimage_futspace.h:52), BASELINE config 1 frame size / distance
    cam = O.Camera(0.98, 0.6, 58, 2.2, 200, 1000, 1.2, 0xFF9090E0)
    frames["live_init_768x1024_d1000"] = O.render(cam, O.default_params(), rgb | 0xFF000000, hgt, 768, 1024)
    np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **frames)
    for k, v in frames.items():
        print(k, v.shape, "sky=%.4f" % (v == 0xFF9090E0).mean(), "colours=%d" % len(np.unique(v)))


if __name__ == "__main__":
    main()
