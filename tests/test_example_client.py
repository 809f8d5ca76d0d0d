"""examples/render_ppm.c: the native C-ABI used from plain C, as INTEGRATION.md section A describes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import SKY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_example(fsb, tmp_path):
    exe = tmp_path / "render_ppm"
    libdir = os.path.dirname(fsb.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "render_ppm.c"),
                           "-L", libdir, "-lfutspace_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    return exe


def test_example_builds_and_fails_loudly_without_a_gpu(fsb, tmp_path):
    import torch
    exe = build_example(fsb, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode != 0 and "no usable sm_100 GPU" in p.stderr     # no CPU fallback
    assert not (tmp_path / "frame.ppm").exists()


@pytest.mark.gpu
def test_example_frame_equals_the_oracle(fsb, oracle, tmp_path, fbm1024):
    exe = build_example(fsb, tmp_path)
    p = subprocess.run([str(exe), "640", "400", "700"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    raw = (tmp_path / "frame.ppm").read_bytes()
    header = b"P6\n640 400\n255\n"
    assert raw.startswith(header)
    rgb = np.frombuffer(raw[len(header):], np.uint8).reshape(400, 640, 3).astype(np.uint32)
    got = (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2]
    col, hgt = fbm1024
    cam = oracle.Camera(np.float32(512.37), np.float32(512.73), 200.0, np.float32(2.2), np.float32(0.3) * np.float32(400), 700.0,
                        np.float32(1.2), SKY)
    want = oracle.render(cam, oracle.default_params(), col, hgt & 0xFF, 400, 640)
    assert np.array_equal(got, want & 0x00FFFFFF)
